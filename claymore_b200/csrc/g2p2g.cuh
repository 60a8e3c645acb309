// g2p2g.cuh -- the fused grid-to-particle / constitutive update / particle-to-grid kernel for sm_100a.
//
// Replaces g2p2g<Partition<1>, GridBuffer, M> (reference Projects/GMPM/mgmpm_kernels.cuh:665-937,
// with fetch_particle_buffer_data :428-462, calculate_contribution_and_store_particle_data :470-663 and
// ParticleBufferImpl::add_advection particle_buffer.cuh:100-135).  Same inputs, same outputs, same
// containers; different machine mapping:
//   * one CTA walks particle blocks (persistent, grid-stride) instead of one CUDA block per particle block;
//   * the 2x2x2 neighbourhood of grid blocks is staged into shared memory with eight 768-byte TMA bulk
//     copies (the three velocity channels of a grid block are contiguous) signalled on an mbarrier,
//     instead of 1536 scalar loads each preceded by a table query (:700-726);
//   * P2G accumulates into a shared arena that has the grid-block layout, so the write-back is eight
//     1-KiB cp.reduce.async.bulk f32-add operations executed by the TMA unit, not 2048 SM-issued global
//     atomics (:910-936);
//   * the 27 neighbour block numbers and source bin offsets are resolved once per block into shared
//     memory instead of two dependent global loads per particle (:761-767, particle_buffer.cuh:101-102).
#pragma once
#include "math3.cuh"

namespace cb200 {

constexpr int kG2P2GThreads = 128;

struct G2P2GArgs {
	Cfg cfg;
	const StepState* state;  // nullable: when set, dt/new_dt/block count come from the device
	float dt, new_dt;
	int block_count;
	int halo_mode;            // 0: all blocks, 1: only halo-marked, 2: only non-halo (MGSP split, mgsp_benchmark.cuh:421-467)
	const char* halo_marks;
	PBuf cur, next;
	Mat mat;
	const int* prev_table;
	const int* table;
	const int* keys;
	const float* grid;
	float* next_grid;
	int* error;  // nullable
};

// compute_dt (utility_funcs.hpp:36-49) evaluated on the device from the reduced max |v|^2
__device__ __forceinline__ float device_compute_dt(const Cfg& cfg, float max_vel_sq, float step_time, float frame_time, float dt_default) {
	float dt = dt_default;
	const float mv = sqrtf(max_vel_sq);
	if(mv > 0.f) dt = fminf(dt, cfg.dx * cfg.cfl / mv);
	if(frame_time > 0.f) dt = fminf(dt, frame_time - step_time);
	return dt;
}

template<int S>
__device__ __forceinline__ int arena_off_x(int X) { return ((X >> 2) << 2) * S + ((X & 3) << 4); }
template<int S>
__device__ __forceinline__ int arena_off_y(int Y) { return ((Y >> 2) << 1) * S + ((Y & 3) << 2); }
template<int S>
__device__ __forceinline__ int arena_off_z(int Z) { return (Z >> 2) * S + (Z & 3); }

template<int MAT>
__global__ void __launch_bounds__(kG2P2GThreads) g2p2g_kernel(const G2P2GArgs a) {
	constexpr int NCH = (MAT == CB200_J_FLUID) ? 4 : (MAT == CB200_FIXED_COROTATED ? 12 : 13);
	constexpr int BINF = (MAT == CB200_J_FLUID) ? 128 : 512;
	constexpr int VS = 192;  // floats per block in the velocity arena (3 channels)
	constexpr int AS = 256;  // floats per block in the accumulation arena (4 channels)
	(void) NCH;

	__shared__ __align__(128) float s_vel[8 * VS];
	__shared__ __align__(128) float s_acc[8 * AS];
	__shared__ int s_nbr[27];
	__shared__ int s_srcbin[27];
	__shared__ __align__(8) uint64_t s_bar;

	const Cfg& cfg = a.cfg;
	const int tid = threadIdx.x;
	float dt = a.dt, new_dt = a.new_dt;
	int nblocks = a.block_count;
	if(a.state) {
		dt = a.state->dt;
		new_dt = device_compute_dt(cfg, a.state->max_vel_sq, a.state->step_time, a.state->frame_time, a.state->dt_default);
		nblocks = a.state->pbc;
	}
	if(tid == 0) {
		mbar_init(&s_bar, 1);
		mbar_fence_init();
	}
	__syncthreads();
	unsigned phase = 0;
	const float dx = cfg.dx, dx_inv = cfg.dx_inv, d_inv = cfg.d_inv;
	const int ppb_mask = cfg.ppb - 1;

	for(int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
		const int bucket_size = a.next.particle_bucket_sizes[blk];
		if(bucket_size == 0) continue;
		if(a.halo_mode) {
			const bool is_halo = a.halo_marks[blk] != 0;
			if((a.halo_mode == 1) != is_halo) continue;
		}
		const int kx = a.keys[3 * blk], ky = a.keys[3 * blk + 1], kz = a.keys[3 * blk + 2];

		// ---- stage the neighbourhood -------------------------------------------------------------
		if(tid < 32) {
			const int lb = tid & 7;
			const int bno = table_query(cfg, a.table, kx + ((lb >> 2) & 1), ky + ((lb >> 1) & 1), kz + (lb & 1));
			const unsigned valid = __ballot_sync(0xffffffffu, tid < 8 && bno >= 0);
			if(tid == 0) mbar_arrive_expect_tx(&s_bar, __popc(valid) * (VS * 4));
			__syncwarp();
			if(tid < 8) {
				if(bno >= 0) {
					tma_load_1d(s_vel + lb * VS, a.grid + (size_t) bno * kGridBlockFloats + 64, VS * 4, &s_bar);
				} else {
					for(int i = 0; i < VS; ++i) s_vel[lb * VS + i] = 0.f;
				}
			}
		} else if(tid < 32 + 27) {
			const int d = tid - 32;
			const int ox = d / 9 - 1, oy = (d / 3) % 3 - 1, oz = d % 3 - 1;
			s_nbr[d] = table_query(cfg, a.table, kx + ox, ky + oy, kz + oz);
			const int pno = table_query(cfg, a.prev_table, kx + ox, ky + oy, kz + oz);
			s_srcbin[d] = pno >= 0 ? a.cur.bin_offsets[pno] : -1;
		}
		{
			float4* acc4 = reinterpret_cast<float4*>(s_acc);
#pragma unroll
			for(int i = tid; i < 8 * AS / 4; i += kG2P2GThreads) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
		__syncthreads();
		mbar_wait(&s_bar, phase);
		phase ^= 1;

		const int dst_bin0 = a.next.bin_offsets[blk];
		const int* __restrict__ bucket = a.next.blockbuckets + ((size_t) blk << cfg.ppb_shift);

		// ---- per-particle work -------------------------------------------------------------------
		for(int pidib = tid; pidib < bucket_size; pidib += kG2P2GThreads) {
			const int advect = __ldg(bucket + pidib);
			const int dir = advect >> cfg.ppb_shift;
			const int src_pidib = advect & ppb_mask;
			const int sbin0 = s_srcbin[dir];
			const float* __restrict__ sbin = a.cur.bins + ((size_t) sbin0 + (src_pidib >> 5)) * BINF + (src_pidib & 31);

			float pos[3] = {__ldg(sbin), __ldg(sbin + 32), __ldg(sbin + 64)};
			int base[3], ab[3];
			float lp[3], w[3][3];
#pragma unroll
			for(int d = 0; d < 3; ++d) {
				base[d] = cell_index(cfg, pos[d]) - 1;
				lp[d] = pos[d] - base[d] * dx;
				bspline_weights(lp[d] * dx_inv, w[d][0], w[d][1], w[d][2]);
				ab[d] = ((base[d] - 1) & 3) + 1;
			}
			// G2P: velocity and APIC matrix (A as in the reference: sum W v (x_i - x_p)^T, column-major A[c + 3d])
			float vel[3] = {0.f, 0.f, 0.f};
			float A[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
			for(int i = 0; i < 3; ++i) {
				const int ox = arena_off_x<VS>(ab[0] + i);
				const float xx = i * dx - lp[0];
#pragma unroll
				for(int j = 0; j < 3; ++j) {
					const int oxy = ox + arena_off_y<VS>(ab[1] + j);
					const float xy = j * dx - lp[1];
					const float wij = w[0][i] * w[1][j];
#pragma unroll
					for(int k = 0; k < 3; ++k) {
						const int o = oxy + arena_off_z<VS>(ab[2] + k);
						const float xz = k * dx - lp[2];
						const float W = wij * w[2][k];
						const float v0 = s_vel[o], v1 = s_vel[o + 64], v2 = s_vel[o + 128];
						const float wv0 = W * v0, wv1 = W * v1, wv2 = W * v2;
						vel[0] += wv0;
						vel[1] += wv1;
						vel[2] += wv2;
						A[0] += wv0 * xx;
						A[1] += wv1 * xx;
						A[2] += wv2 * xx;
						A[3] += wv0 * xy;
						A[4] += wv1 * xy;
						A[5] += wv2 * xy;
						A[6] += wv0 * xz;
						A[7] += wv1 * xz;
						A[8] += wv2 * xz;
					}
				}
			}
#pragma unroll
			for(int d = 0; d < 3; ++d) pos[d] += vel[d] * dt;

			float contrib[9];
			float* __restrict__ dbin = a.next.bins + ((size_t) dst_bin0 + (pidib >> 5)) * BINF + (pidib & 31);
			if constexpr(MAT == CB200_J_FLUID) {
				float J = __ldg(sbin + 96);
				J += (A[0] + A[4] + A[8]) * dt * d_inv * J;
				if(J < 0.1f) J = 0.1f;
				const float voln = J * a.mat.volume;
				const float pressure = a.mat.bulk * (powf(J, -a.mat.gamma) - 1.f);
				const float vs = d_inv * a.mat.viscosity;
				contrib[0] = ((A[0] + A[0]) * vs - pressure) * voln;
				contrib[1] = (A[1] + A[3]) * vs * voln;
				contrib[2] = (A[2] + A[6]) * vs * voln;
				contrib[3] = contrib[1];
				contrib[4] = ((A[4] + A[4]) * vs - pressure) * voln;
				contrib[5] = (A[5] + A[7]) * vs * voln;
				contrib[6] = contrib[2];
				contrib[7] = contrib[5];
				contrib[8] = ((A[8] + A[8]) * vs - pressure) * voln;
				dbin[0] = pos[0];
				dbin[32] = pos[1];
				dbin[64] = pos[2];
				dbin[96] = J;
			} else {
				float Fo[9], F[9], G[9];
#pragma unroll
				for(int d = 0; d < 9; ++d) Fo[d] = __ldg(sbin + (3 + d) * 32);
				const float sc = dt * d_inv;
#pragma unroll
				for(int d = 0; d < 9; ++d) G[d] = A[d] * sc + ((d & 3) ? 0.f : 1.f);
#pragma unroll
				for(int c = 0; c < 3; ++c)
#pragma unroll
					for(int r = 0; r < 3; ++r) F[r + 3 * c] = G[r] * Fo[3 * c] + G[r + 3] * Fo[3 * c + 1] + G[r + 6] * Fo[3 * c + 2];
				dbin[0] = pos[0];
				dbin[32] = pos[1];
				dbin[64] = pos[2];
				if constexpr(MAT == CB200_FIXED_COROTATED) {
#pragma unroll
					for(int d = 0; d < 9; ++d) dbin[(3 + d) * 32] = F[d];
					stress_fixed_corotated(a.mat, F, contrib);
				} else {
					float log_jp = __ldg(sbin + 12 * 32);
					if constexpr(MAT == CB200_SAND) stress_sand(a.mat, F, contrib, log_jp);
					else stress_nacc(a.mat, F, contrib, log_jp);
#pragma unroll
					for(int d = 0; d < 9; ++d) dbin[(3 + d) * 32] = F[d];
					dbin[12 * 32] = log_jp;
				}
			}
			const float mass = a.mat.mass;
#pragma unroll
			for(int d = 0; d < 9; ++d) contrib[d] = (A[d] * mass - contrib[d] * new_dt) * d_inv;

			// ---- re-bucket (add_advection) -------------------------------------------------------
			int nb[3], rel[3], cell[3];
			bool far = false;
#pragma unroll
			for(int d = 0; d < 3; ++d) {
				nb[d] = cell_index(cfg, pos[d]) - 1;
				lp[d] = pos[d] - nb[d] * dx;
				cell[d] = nb[d] - 1;
			}
			rel[0] = (cell[0] >> 2) - kx;
			rel[1] = (cell[1] >> 2) - ky;
			rel[2] = (cell[2] >> 2) - kz;
#pragma unroll
			for(int d = 0; d < 3; ++d) far |= (rel[d] < -1) | (rel[d] > 1);
			if(!far) {
				const int bno = s_nbr[(rel[0] + 1) * 9 + (rel[1] + 1) * 3 + rel[2] + 1];
				if(bno >= 0) {
					const int dirtag = (1 - rel[0]) * 9 + (1 - rel[1]) * 3 + (1 - rel[2]);
					const int cellno = ((cell[0] & 3) << 4) | ((cell[1] & 3) << 2) | (cell[2] & 3);
					int* cnt = a.next.cell_particle_counts + (size_t) bno * kBlockVol + cellno;
					const int slot = atomicAdd(cnt, 1);
					if(slot >= cfg.max_ppc) {
						atomicSub(cnt, 1);
						if(a.error) atomicOr(a.error, kErrCellOverflow);
					} else {
						a.next.cellbuckets[((size_t) bno << cfg.ppb_shift) + (cellno << cfg.ppc_shift) + slot] = (dirtag << cfg.ppb_shift) | pidib;
					}
				} else if(a.error) {
					atomicOr(a.error, kErrLostParticle);
				}
			} else if(a.error) {
				atomicOr(a.error, kErrLostParticle);
			}

			// ---- P2G into the shared arena -------------------------------------------------------
			bool oob = false;
#pragma unroll
			for(int d = 0; d < 3; ++d) {
				bspline_weights(lp[d] * dx_inv, w[d][0], w[d][1], w[d][2]);
				ab[d] = ab[d] + (nb[d] - base[d]);
				oob |= (ab[d] < 0) | (ab[d] > 5);
			}
			if(oob) {  // moved more than one cell: the reference drops the contribution (mgmpm_kernels.cuh:881-885)
				if(a.error) atomicOr(a.error, kErrLostParticle);
				continue;
			}
#pragma unroll
			for(int i = 0; i < 3; ++i) {
				const int ox = arena_off_x<AS>(ab[0] + i);
				const float xx = i * dx - lp[0];
#pragma unroll
				for(int j = 0; j < 3; ++j) {
					const int oxy = ox + arena_off_y<AS>(ab[1] + j);
					const float xy = j * dx - lp[1];
					const float wij = w[0][i] * w[1][j];
#pragma unroll
					for(int k = 0; k < 3; ++k) {
						const int o = oxy + arena_off_z<AS>(ab[2] + k);
						const float xz = k * dx - lp[2];
						const float W = wij * w[2][k];
						const float wm = mass * W;
						atomicAdd(&s_acc[o], wm);
						atomicAdd(&s_acc[o + 64], wm * vel[0] + (contrib[0] * xx + contrib[3] * xy + contrib[6] * xz) * W);
						atomicAdd(&s_acc[o + 128], wm * vel[1] + (contrib[1] * xx + contrib[4] * xy + contrib[7] * xz) * W);
						atomicAdd(&s_acc[o + 192], wm * vel[2] + (contrib[2] * xx + contrib[5] * xy + contrib[8] * xz) * W);
					}
				}
			}
		}

		// ---- arena -> next grid: eight 1-KiB bulk add-reductions ----------------------------------
		fence_proxy_async();
		__syncthreads();
		if(tid < 8) {
			const int bno = table_query(cfg, a.table, kx + ((tid >> 2) & 1), ky + ((tid >> 1) & 1), kz + (tid & 1));
			if(bno >= 0) tma_reduce_add_f32(a.next_grid + (size_t) bno * kGridBlockFloats, s_acc + tid * AS, AS * 4);
			tma_commit();
			tma_wait_read<0>();
		}
		__syncthreads();
	}
	if(threadIdx.x < 8) tma_wait_all<0>();
}

}  // namespace cb200
