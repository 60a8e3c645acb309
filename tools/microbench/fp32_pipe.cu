// Micro-benchmark: issue rate and latency of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32_pipe tools/microbench/fp32_pipe.cu && /tmp/fp32_pipe
// Prints FP32 FMA lanes per clock per SM for: independent-chain throughput at 4 / 8 / 16 warps per SM sub-partition, and the
// dependent-chain latency of one warp.  Decides whether packing the separable-weight arithmetic of g2p2g can pay.
#include <cstdio>
#include <cuda_runtime.h>

template<int MODE, int ILP>
__global__ void __launch_bounds__(1024) kern(float* out, int iters, float s, long long* cycles) {
	float2 acc[ILP];
#pragma unroll
	for(int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
	const float2 m = make_float2(s, s * 0.999f);
	float2 w[ILP];  // loop-invariant multiplicands: every FMA below reads three registers, like the accumulations of g2p2g
#pragma unroll
	for(int i = 0; i < ILP; ++i) w[i] = make_float2(1e-7f * (threadIdx.x + i), 2e-7f * (threadIdx.x + i));
	const long long t0 = clock64();
	for(int it = 0; it < iters; ++it) {
#pragma unroll
		for(int i = 0; i < ILP; ++i) {
			if(MODE == 0) {  // scalar: two FFMA
				acc[i].x = fmaf(w[i].x, m.x, acc[i].x);
				acc[i].y = fmaf(w[i].y, m.y, acc[i].y);
			} else if(MODE == 1) {  // packed, pair operands
				acc[i] = __ffma2_rn(w[i], m, acc[i]);
			} else {  // packed, broadcast scalar operand
				acc[i] = __ffma2_rn(w[i], make_float2(s, s), acc[i]);
			}
		}
	}
	const long long t1 = clock64();
	float r = 0.f;
#pragma unroll
	for(int i = 0; i < ILP; ++i) r += acc[i].x + acc[i].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
	if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template<int MODE, int ILP>
void run(const char* name, int threads, int sms) {
	float* out;
	long long* cyc;
	cudaMalloc(&out, sizeof(float) * threads * sms);
	cudaMalloc(&cyc, sizeof(long long) * sms);
	const int iters = 4096;
	kern<MODE, ILP><<<sms, threads>>>(out, 16, 0.9999f, cyc);
	kern<MODE, ILP><<<sms, threads>>>(out, iters, 0.9999f, cyc);
	cudaDeviceSynchronize();
	long long h[1024];
	cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
	double avg = 0;
	for(int i = 0; i < sms; ++i) avg += (double) h[i];
	avg /= sms;
	const double fmas = (double) iters * ILP * 2 * threads;  // FP32 FMA lane-operations per SM
	const double inst = (double) iters * ILP * (MODE == 0 ? 2 : 1) * (threads / 32);
	printf("%-34s threads/SM %4d ILP %2d: %7.1f FMA lanes/clk/SM, %5.2f warp-inst/clk/SM, %6.2f clk per dependent step\n", name, threads, ILP, fmas / avg, inst / avg, avg / iters);
	cudaFree(out);
	cudaFree(cyc);
}

int main() {
	int sms = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	printf("SMs %d\n", sms);
	for(int threads : {128, 512, 1024}) {
		if(threads == 128) {
			run<0, 8>("FFMA  scalar", threads, sms);
			run<1, 8>("FFMA2 pair operands", threads, sms);
			run<2, 8>("FFMA2 broadcast operand", threads, sms);
		} else if(threads == 512) {
			run<0, 8>("FFMA  scalar", threads, sms);
			run<1, 8>("FFMA2 pair operands", threads, sms);
			run<2, 8>("FFMA2 broadcast operand", threads, sms);
		} else {
			run<0, 8>("FFMA  scalar", threads, sms);
			run<1, 8>("FFMA2 pair operands", threads, sms);
			run<2, 8>("FFMA2 broadcast operand", threads, sms);
		}
	}
	// latency: one warp per SM, one dependent chain
	run<0, 1>("FFMA  scalar (latency, 1 warp)", 32, sms);
	run<1, 1>("FFMA2 pair   (latency, 1 warp)", 32, sms);
	run<2, 1>("FFMA2 bcast  (latency, 1 warp)", 32, sms);
	return 0;
}
