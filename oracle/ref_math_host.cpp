// ref_math_host.cpp -- compiles the REFERENCE's own 3x3 SVD and constitutive models for the host.
// TEST INFRASTRUCTURE ONLY.  Nothing is copied: the reference headers are included where they lie
// (/root/reference/Library/MnBase/Math/Matrix/svd.cuh, Projects/GMPM/constitutive_models.cuh);
// this file only supplies host stand-ins for the CUDA keywords/intrinsics they use.
//   - QR_CUH is pre-defined by the build recipe so the broken, unused 2-D overload in qr.cuh:55
//     is never parsed (SURVEY.md "Read this first" #1); the 2-D polar helper is declared, not defined.
// Output: oracle/_ref/libclaymore_ref_math.so (git-ignored, built by oracle/build_ref.sh).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <limits>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __frsqrt_rn(float a) { return (float) (1.0 / std::sqrt((double) a)); }
static inline double __fadd_rn(double a, double b) { return a + b; }
static inline double __fsub_rn(double a, double b) { return a - b; }
static inline double __frsqrt_rn(double a) { return 1.0 / std::sqrt(a); }
using std::max;
using std::min;

#include <MnBase/Math/Matrix/Givens.cuh>
namespace mn { namespace math {
template<typename T> void polar_decomposition(const std::array<T, 4>& a, GivensRotation<T>& r, std::array<T, 4>& s);
}}
#include <constitutive_models.cuh>

extern "C" {
void ref_svd3(const float* F, float* U, float* S, float* V) {
	std::array<float, 9> f, u, v;
	std::array<float, 3> s;
	for(int i = 0; i < 9; ++i) f[i] = F[i];
	mn::math::svd<float, 3>(f, u, s, v);
	for(int i = 0; i < 9; ++i) { U[i] = u[i]; V[i] = v[i]; }
	for(int i = 0; i < 3; ++i) S[i] = s[i];
}
// params: volume, mu, lambda, then ComputeStressIntermediate fields {bm, xi, beta, msqr, log_jp, cohesion, yield_surface, hardening_on, volume_correction}
void ref_compute_stress(int material, const float* params, float* F, float* PF, float* log_jp) {
	std::array<float, 9> f, pf;
	for(int i = 0; i < 9; ++i) f[i] = F[i];
	mn::ComputeStressIntermediate<float> d = {};
	d.bm = params[3]; d.xi = params[4]; d.beta = params[5]; d.msqr = params[6]; d.log_jp = *log_jp;
	d.cohesion = params[8]; d.yield_surface = params[9]; d.hardening_on = params[10] != 0.f; d.volume_correction = params[11] != 0.f;
	switch(material) {
		case 1: mn::compute_stress<float, mn::MaterialE::FIXED_COROTATED>(params[0], params[1], params[2], f, pf, d); break;
		case 2: mn::compute_stress<float, mn::MaterialE::SAND>(params[0], params[1], params[2], f, pf, d); break;
		case 3: mn::compute_stress<float, mn::MaterialE::NACC>(params[0], params[1], params[2], f, pf, d); break;
		default: return;
	}
	for(int i = 0; i < 9; ++i) { F[i] = f[i]; PF[i] = pf[i]; }
	*log_jp = d.log_jp;
}
}
