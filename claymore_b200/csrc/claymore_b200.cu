// claymore_b200.cu -- single translation unit of libclaymore_b200.so (kernels are defined in headers).
#include "capi.cu"
#include "engine.cu"
