// ref_gpu_driver.cu -- runs the REFERENCE's own GMPM kernels (compiled unmodified for sm_100a) behind a small C ABI.
// TEST INFRASTRUCTURE ONLY.  The reference's headers are included where they lie under /root/reference
// (Projects/GMPM/{mgmpm_kernels,particle_buffer,grid_buffer,hash_table}.cuh and what they include); nothing is copied.
// What is written here is only the host loop: it issues the reference's kernels in the order, with the launch shapes,
// memsets, thrust scans and device->host counter copies of GmpmSimulator::initial_setup / main_loop
// (Projects/GMPM/gmpm_simulator.cuh:324-580, 637-781), without its fmt prints, file output and resize policy.
// The reference's own driver (gmpm.cu) cannot be built here: it needs cxxopts, rapidjson, fmt and partio downloads.
//
// Build recipe: oracle/build_ref.sh (one .so per DOMAIN_BITS: the reference's sizes are compile-time constants; the recipe
// pre-includes a sed-generated settings header from the git-ignored build directory and pre-defines QR_CUH, see there).
// Two uses: (1) parity pin -- outputs of the reference itself for the oracle and the CUDA path to be compared with;
//           (2) the "reference claymore on B200" particle-steps/s denominator of BASELINE.md section 2a.
#include <MnBase/Math/Matrix/Givens.cuh>
namespace mn { namespace math {
template<typename T> __host__ __device__ void polar_decomposition(const std::array<T, 4>& a, GivensRotation<T>& r, std::array<T, 4>& s);
}}
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

#include <cstdio>
#include <vector>

#include <grid_buffer.cuh>
#include <hash_table.cuh>
#include <mgmpm_kernels.cuh>
#include <particle_buffer.cuh>

// ---- optional: swap single kernels of the reference's loop for libclaymore_b200's (drop-in test) ------------------------------
// The product library is resolved at run time (dlopen), so this checker has no link-time dependency on it; the calls go through
// include/claymore_b200_adapter.cuh -- the binding a maintainer of the reference would add -- with the reference's own containers.
#include <dlfcn.h>

#include <claymore_b200.h>
#define B200_FUNCS(X)                                                                                                                                                      \
	X(cb200_g2p2g) X(cb200_update_grid_velocity_query_max) X(cb200_clear_grid) X(cb200_cell_bucket_to_block) X(cb200_compute_bin_capacity) X(cb200_init_adv_bucket)       \
	X(cb200_register_neighbor_blocks) X(cb200_register_exterior_blocks) X(cb200_mark_active_grid_blocks) X(cb200_mark_active_particle_blocks) X(cb200_exclusive_scan_inverse) \
	X(cb200_update_partition) X(cb200_update_buckets) X(cb200_copy_selected_grid_blocks) X(cb200_activate_blocks) X(cb200_build_particle_cell_buckets) X(cb200_array_to_buffer) \
	X(cb200_rasterize) X(cb200_retrieve_particle_buffer)
#define X(n) static decltype(&::n) g_b200_##n = nullptr;
B200_FUNCS(X)
#undef X
#define CB200_ADAPTER_CALL(name) (*g_b200_##name)
#include <claymore_b200_adapter.cuh>
static int g_swap = 0;  // bit 0: g2p2g, bit 1: update_grid_velocity_query_max, bit 2: partition / bucket / grid-carry group, bit 3: init kernels

using namespace mn;

namespace {
struct DevAlloc {
	void* allocate(std::size_t bytes) {
		void* p = nullptr;
		check_cuda_errors(cudaMalloc(&p, bytes));
		return p;
	}
	void deallocate(void* p, std::size_t) { check_cuda_errors(cudaFree(p)); }
};

// export of all particle channels with the traversal of retrieve_particle_buffer (positions only in the reference)
template<typename PB>
__global__ void export_state(int nch, Partition<1> part, Partition<1> prev, PB pb, PB next_pb, float* out, int* parcount) {
	const int cnt = next_pb.particle_bucket_sizes[blockIdx.x];
	const ivec3 blockid = part.active_keys[blockIdx.x];
	const int* bucket = next_pb.blockbuckets + (size_t) blockIdx.x * config::G_PARTICLE_NUM_PER_BLOCK;
	for(int i = threadIdx.x; i < cnt; i += blockDim.x) {
		const int advect = bucket[i];
		ivec3 src;
		dir_components(advect / config::G_PARTICLE_NUM_PER_BLOCK, src.data_arr());
		src += blockid;
		const int sp = advect % config::G_PARTICLE_NUM_PER_BLOCK;
		const int sno = prev.query(src);
		const float* bin = reinterpret_cast<const float*>(pb.handle.ptr) + ((size_t) pb.bin_offsets[sno] + sp / 32) * (nch == 4 ? 128 : 512) + sp % 32;
		const int o = atomicAdd(parcount, 1);
		for(int c = 0; c < nch; ++c) out[(size_t) o * nch + c] = bin[c * 32];
	}
}

struct SimBase {
	virtual ~SimBase() {}
	virtual int init_model(const float* pos, int n, const float* v0, const float* params) = 0;
	virtual int setup() = 0;
	virtual int step(int n) = 0;
	virtual void counts(int* out) = 0;
	virtual int keys(int* out) = 0;
	virtual int grid(float* out) = 0;
	virtual int state(int model, float* out) = 0;
	virtual float dt_now() = 0;
};

template<MaterialE M>
struct Sim : SimBase {
	using PB = ParticleBuffer<M>;
	static constexpr int NCH = M == MaterialE::J_FLUID ? 4 : (M == MaterialE::FIXED_COROTATED ? 12 : 13);
	float dt_default, dt, next_dt, max_vel = 0.f;
	int rollid = 0;
	cudaStream_t st = nullptr;
	std::vector<GridBuffer> grids;
	std::vector<Partition<1>> parts;
	std::vector<PB> bins[2];
	std::vector<float*> d_pos;
	std::vector<int> counts_, bincount;
	std::vector<std::array<float, 3>> v0s;
	int *marks = nullptr, *dest = nullptr, *sources = nullptr, *bin_sizes = nullptr;
	float* d_max_vel = nullptr;
	int pbc = 0, nbc = 0, ebc = 0;

	explicit Sim(float dtd) : dt_default(dtd), dt(dtd), next_dt(dtd) {
		check_cuda_errors(cudaStreamCreate(&st));
		const size_t mb = config::G_MAX_ACTIVE_BLOCK;
		for(int i = 0; i < 2; ++i) {
			grids.emplace_back(DevAlloc {});
			parts.emplace_back(DevAlloc {}, (int) mb);
		}
		check_cuda_errors(cudaMalloc(&marks, sizeof(int) * (mb + 2)));
		check_cuda_errors(cudaMalloc(&dest, sizeof(int) * (mb + 2)));
		check_cuda_errors(cudaMalloc(&sources, sizeof(int) * (mb + 2)));
		check_cuda_errors(cudaMalloc(&bin_sizes, sizeof(int) * (mb + 2)));
		check_cuda_errors(cudaMalloc(&d_max_vel, sizeof(float)));
	}
	int init_model(const float* pos, int n, const float* v0, const float* params) override {
		for(int i = 0; i < 2; ++i) {
			bins[i].emplace_back(DevAlloc {}, (std::size_t) n / config::G_BIN_CAPACITY + config::G_MAX_ACTIVE_BLOCK);
			bins[i].back().reserve_buckets(DevAlloc {}, config::G_MAX_ACTIVE_BLOCK);
			set_params(bins[i].back(), params);
		}
		float* d = nullptr;
		check_cuda_errors(cudaMalloc(&d, sizeof(float) * 3 * (size_t) n));
		check_cuda_errors(cudaMemcpy(d, pos, sizeof(float) * 3 * (size_t) n, cudaMemcpyHostToDevice));
		d_pos.push_back(d);
		counts_.push_back(n);
		bincount.push_back(0);
		v0s.push_back({v0[0], v0[1], v0[2]});
		return (int) d_pos.size() - 1;
	}
	// params: rho, volume, E, nu (elastic) or rho, volume, bulk, gamma, viscosity (fluid); nullptr keeps the defaults
	static void set_params(PB& pb, const float* p) {
		if(!p) return;
		if constexpr(M == MaterialE::J_FLUID) pb.update_parameters(p[0], p[1], p[2], p[3], p[4]);
		else if constexpr(M == MaterialE::FIXED_COROTATED) pb.update_parameters(p[0], p[1], p[2], p[3]);
		else if constexpr(M == MaterialE::NACC) pb.update_parameters(p[0], p[1], p[2], p[3], p[4], p[5]);
		else {  // SAND has no update_parameters in the reference: set the same fields by hand
			pb.rho = p[0];
			pb.volume = p[1];
			pb.mass = p[1] * p[0];
			pb.lambda = p[2] * p[3] / ((1 + p[3]) * (1 - 2 * p[3]));
			pb.mu = p[2] / (2 * (1 + p[3]));
		}
	}
	ParticleArray parray(int m) {
		ParticleArray a {};
		a.handle.ptr = d_pos[m];
		return a;
	}
	void scan(int count, const int* in, int* out) { thrust::exclusive_scan(thrust::cuda::par.on(st), in, in + count, out); }
	void fetch(int* host, const int* dev) {
		check_cuda_errors(cudaMemcpyAsync(host, dev, sizeof(int), cudaMemcpyDefault, st));
		check_cuda_errors(cudaStreamSynchronize(st));
	}

	// GmpmSimulator::initial_setup, gmpm_simulator.cuh:637-781
	int setup() override {
		const int R = rollid, Rn = R ^ 1;
		const int nm = (int) d_pos.size();
		float mv = 0.f;
		for(auto& v : v0s) mv = fmaxf(mv, sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
		dt = compute_dt(mv, Duration::zero(), Duration(1e30f), Duration(dt_default)).count();
		for(int m = 0; m < nm; ++m) {
			if(g_swap & 8) check_cuda_errors(b200::launch_on(st, {(counts_[m] + 255) / 256, 256}, b200::activate_blocks, (uint32_t) counts_[m], parray(m), parts[Rn]));
			else activate_blocks<<<(counts_[m] + 255) / 256, 256, 0, st>>>((uint32_t) counts_[m], parray(m), parts[Rn]);
		}
		fetch(&pbc, parts[Rn].count);
		for(int m = 0; m < nm; ++m) {
			if(g_swap & 8) check_cuda_errors(b200::launch_on(st, {(counts_[m] + 255) / 256, 256}, b200::build_particle_cell_buckets, (uint32_t) counts_[m], parray(m), bins[R][m], parts[Rn]));
			else build_particle_cell_buckets<<<(counts_[m] + 255) / 256, 256, 0, st>>>((uint32_t) counts_[m], parray(m), bins[R][m], parts[Rn]);
		}
		for(int m = 0; m < nm; ++m) {
			PB& pb = bins[R][m];
			check_cuda_errors(cudaMemsetAsync(pb.particle_bucket_sizes, 0, sizeof(int) * (pbc + 1), st));
			if(g_swap & 8) {
				check_cuda_errors(b200::launch_on(st, {pbc, config::G_BLOCKVOLUME}, b200::cell_bucket_to_block, (const int*) pb.cell_particle_counts, (const int*) pb.cellbuckets, pb.particle_bucket_sizes, pb.blockbuckets));
				check_cuda_errors(b200::launch_on(st, {pbc / 128 + 1, 128}, b200::compute_bin_capacity, (uint32_t) (pbc + 1), (const int*) pb.particle_bucket_sizes, bin_sizes));
			} else {
				cell_bucket_to_block<<<pbc, config::G_BLOCKVOLUME, 0, st>>>(pb.cell_particle_counts, pb.cellbuckets, pb.particle_bucket_sizes, pb.blockbuckets);
				compute_bin_capacity<<<pbc / 128 + 1, 128, 0, st>>>((uint32_t) (pbc + 1), (const int*) pb.particle_bucket_sizes, bin_sizes);
			}
			scan(pbc + 1, bin_sizes, pb.bin_offsets);
			fetch(&bincount[m], pb.bin_offsets + pbc);
			if(g_swap & 8) check_cuda_errors(b200::launch_on(st, {pbc, 128}, b200::array_to_buffer, parray(m), pb));
			else array_to_buffer<<<pbc, 128, 0, st>>>(parray(m), pb);
		}
		register_neighbor_blocks<<<(pbc + 127) / 128, 128, 0, st>>>((uint32_t) pbc, parts[Rn]);
		fetch(&nbc, parts[Rn].count);
		register_exterior_blocks<<<(pbc + 127) / 128, 128, 0, st>>>((uint32_t) pbc, parts[Rn]);
		fetch(&ebc, parts[Rn].count);
		if(ebc > (int) config::G_MAX_ACTIVE_BLOCK) return -1;
		parts[Rn].copy_to(parts[R], ebc, st);
		check_cuda_errors(cudaMemcpyAsync(parts[R].active_keys, parts[Rn].active_keys, sizeof(ivec3) * ebc, cudaMemcpyDefault, st));
		for(int m = 0; m < nm; ++m) bins[R][m].copy_to(bins[Rn][m], pbc, st);
		check_cuda_errors(cudaStreamSynchronize(st));
		clear_grid<<<nbc, config::G_BLOCKVOLUME, 0, st>>>(grids[0]);
		for(int m = 0; m < nm; ++m) {
			if(g_swap & 8) {
				check_cuda_errors(b200::launch_on(st, {(counts_[m] + 255) / 256, 256}, b200::rasterize, (uint32_t) counts_[m], (const ParticleArray) parray(m), grids[0], (const Partition<1>) parts[R], Duration(dt), bins[R][m].mass, v0s[m]));
				check_cuda_errors(b200::launch_on(st, {pbc, 128}, b200::init_adv_bucket, (const int*) bins[Rn][m].particle_bucket_sizes, bins[Rn][m].blockbuckets));
			} else {
				rasterize<<<(counts_[m] + 255) / 256, 256, 0, st>>>((uint32_t) counts_[m], parray(m), grids[0], parts[R], Duration(dt), bins[R][m].mass, v0s[m]);
				init_adv_bucket<<<pbc, 128, 0, st>>>((const int*) bins[Rn][m].particle_bucket_sizes, bins[Rn][m].blockbuckets);
			}
		}
		check_cuda_errors(cudaStreamSynchronize(st));
		return 0;
	}

	// one pass of the inner loop of GmpmSimulator::main_loop, gmpm_simulator.cuh:324-580
	int substep() {
		const int R = rollid, Rn = R ^ 1;
		const int nm = (int) d_pos.size();
		check_cuda_errors(cudaMemsetAsync(d_max_vel, 0, sizeof(float), st));
		if(g_swap & 2) check_cuda_errors(b200::launch_on(st, {(nbc + config::G_NUM_GRID_BLOCKS_PER_CUDA_BLOCK - 1) / config::G_NUM_GRID_BLOCKS_PER_CUDA_BLOCK, config::G_NUM_WARPS_PER_CUDA_BLOCK * config::CUDA_WARP_SIZE * config::G_NUM_WARPS_PER_GRID_BLOCK}, b200::update_grid_velocity_query_max, (uint32_t) nbc, grids[0], parts[R], Duration(dt), d_max_vel));
		else update_grid_velocity_query_max<<<(nbc + config::G_NUM_GRID_BLOCKS_PER_CUDA_BLOCK - 1) / config::G_NUM_GRID_BLOCKS_PER_CUDA_BLOCK, config::G_NUM_WARPS_PER_CUDA_BLOCK * config::CUDA_WARP_SIZE * config::G_NUM_WARPS_PER_GRID_BLOCK, 0, st>>>((uint32_t) nbc, grids[0], parts[R], Duration(dt), d_max_vel);
		check_cuda_errors(cudaMemcpyAsync(&max_vel, d_max_vel, sizeof(float), cudaMemcpyDefault, st));
		check_cuda_errors(cudaStreamSynchronize(st));
		if(std::isinf(max_vel)) return -2;
		max_vel = std::sqrt(max_vel);
		next_dt = compute_dt(max_vel, Duration::zero(), Duration(1e30f), Duration(dt_default)).count();
		clear_grid<<<nbc, config::G_BLOCKVOLUME, 0, st>>>(grids[1]);
		for(int m = 0; m < nm; ++m) {
			check_cuda_errors(cudaMemsetAsync(bins[Rn][m].cell_particle_counts, 0, sizeof(int) * (size_t) ebc * config::G_BLOCKVOLUME, st));
			if(g_swap & 1) check_cuda_errors(b200::launch_on(st, {pbc, config::G_PARTICLE_BATCH_CAPACITY}, b200::g2p2g, Duration(dt), Duration(next_dt), (const PB) bins[R][m], bins[Rn][m], (const Partition<1>) parts[Rn], parts[R], (const GridBuffer) grids[0], grids[1]));
			else g2p2g<<<pbc, config::G_PARTICLE_BATCH_CAPACITY, 0, st>>>(Duration(dt), Duration(next_dt), (const PB) bins[R][m], bins[Rn][m], (const Partition<1>) parts[Rn], parts[R], (const GridBuffer) grids[0], grids[1]);
		}
		check_cuda_errors(cudaStreamSynchronize(st));
		for(int m = 0; m < nm; ++m) {
			PB& pb = bins[Rn][m];
			check_cuda_errors(cudaMemsetAsync(pb.particle_bucket_sizes, 0, sizeof(int) * (ebc + 1), st));
			if(g_swap & 4) check_cuda_errors(b200::launch_on(st, {ebc, config::G_BLOCKVOLUME}, b200::cell_bucket_to_block, (const int*) pb.cell_particle_counts, (const int*) pb.cellbuckets, pb.particle_bucket_sizes, pb.blockbuckets));
			else cell_bucket_to_block<<<ebc, config::G_BLOCKVOLUME, 0, st>>>(pb.cell_particle_counts, pb.cellbuckets, pb.particle_bucket_sizes, pb.blockbuckets);
		}
		check_cuda_errors(cudaMemsetAsync(marks, 0, sizeof(int) * nbc, st));
		if(g_swap & 4) check_cuda_errors(b200::launch_on(st, {(nbc * config::G_BLOCKVOLUME + 127) / 128, 128}, b200::mark_active_grid_blocks, (uint32_t) nbc, (const GridBuffer) grids[1], marks));
		else mark_active_grid_blocks<<<(nbc * config::G_BLOCKVOLUME + 127) / 128, 128, 0, st>>>((uint32_t) nbc, (const GridBuffer) grids[1], marks);
		check_cuda_errors(cudaMemsetAsync(sources, 0, sizeof(int) * (ebc + 1), st));
		for(int m = 0; m < nm; ++m) {
			if(g_swap & 4) check_cuda_errors(b200::launch_on(st, {ebc / 128 + 1, 128}, b200::mark_active_particle_blocks, (uint32_t) (ebc + 1), (const int*) bins[Rn][m].particle_bucket_sizes, sources));
			else mark_active_particle_blocks<<<ebc / 128 + 1, 128, 0, st>>>((uint32_t) (ebc + 1), (const int*) bins[Rn][m].particle_bucket_sizes, sources);
		}
		scan(ebc + 1, sources, dest);
		check_cuda_errors(cudaMemcpyAsync(parts[Rn].count, dest + ebc, sizeof(int), cudaMemcpyDefault, st));
		check_cuda_errors(cudaMemcpyAsync(&pbc, dest + ebc, sizeof(int), cudaMemcpyDefault, st));
		if(g_swap & 4) check_cuda_errors(b200::launch_on(st, {(ebc + 255) / 256, 256}, b200::exclusive_scan_inverse, ebc, (const int*) dest, sources));
		else exclusive_scan_inverse<<<(ebc + 255) / 256, 256, 0, st>>>(ebc, (const int*) dest, sources);
		parts[Rn].reset_table(st);
		check_cuda_errors(cudaStreamSynchronize(st));
		if(g_swap & 4) check_cuda_errors(b200::launch_on(st, {(pbc + 127) / 128, 128}, b200::update_partition, (uint32_t) pbc, (const int*) sources, (const Partition<1>) parts[R], parts[Rn]));
		else update_partition<<<(pbc + 127) / 128, 128, 0, st>>>((uint32_t) pbc, (const int*) sources, (const Partition<1>) parts[R], parts[Rn]);
		for(int m = 0; m < nm; ++m) {
			if(g_swap & 4) {
				check_cuda_errors(b200::launch_on(st, {pbc, 128}, b200::update_buckets, (uint32_t) pbc, (const int*) sources, (const PB) bins[Rn][m], bins[R][m]));
				check_cuda_errors(b200::launch_on(st, {pbc / 128 + 1, 128}, b200::compute_bin_capacity, (uint32_t) (pbc + 1), (const int*) bins[R][m].particle_bucket_sizes, bin_sizes));
			} else {
				update_buckets<<<pbc, 128, 0, st>>>((uint32_t) pbc, (const int*) sources, (const PB) bins[Rn][m], bins[R][m]);
				compute_bin_capacity<<<pbc / 128 + 1, 128, 0, st>>>((uint32_t) (pbc + 1), (const int*) bins[R][m].particle_bucket_sizes, bin_sizes);
			}
			scan(pbc + 1, bin_sizes, bins[R][m].bin_offsets);
			fetch(&bincount[m], bins[R][m].bin_offsets + pbc);
		}
		if(g_swap & 4) check_cuda_errors(b200::launch_on(st, {(pbc + 127) / 128, 128}, b200::register_neighbor_blocks, (uint32_t) pbc, parts[Rn]));
		else register_neighbor_blocks<<<(pbc + 127) / 128, 128, 0, st>>>((uint32_t) pbc, parts[Rn]);
		const int prev_nbc = nbc;
		fetch(&nbc, parts[Rn].count);
		if(g_swap & 4) {
			check_cuda_errors(b200::launch_on(st, {ebc, config::G_BLOCKVOLUME}, b200::clear_grid, grids[0]));
			check_cuda_errors(b200::launch_on(st, {prev_nbc, config::G_BLOCKVOLUME}, b200::copy_selected_grid_blocks, (const ivec3*) parts[R].active_keys, (const Partition<1>) parts[Rn], (const int*) marks, grids[1], grids[0]));
		} else {
			clear_grid<<<ebc, config::G_BLOCKVOLUME, 0, st>>>(grids[0]);
			copy_selected_grid_blocks<<<prev_nbc, config::G_BLOCKVOLUME, 0, st>>>((const ivec3*) parts[R].active_keys, (const Partition<1>) parts[Rn], (const int*) marks, grids[1], grids[0]);
		}
		check_cuda_errors(cudaStreamSynchronize(st));
		if(g_swap & 4) check_cuda_errors(b200::launch_on(st, {(pbc + 127) / 128, 128}, b200::register_exterior_blocks, (uint32_t) pbc, parts[Rn]));
		else register_exterior_blocks<<<(pbc + 127) / 128, 128, 0, st>>>((uint32_t) pbc, parts[Rn]);
		fetch(&ebc, parts[Rn].count);
		if(ebc > (int) config::G_MAX_ACTIVE_BLOCK) return -1;
		rollid = Rn;
		dt = next_dt;
		return 0;
	}
	int step(int n) override {
		for(int i = 0; i < n; ++i) {
			const int e = substep();
			if(e) return e;
		}
		return (int) cudaGetLastError();
	}
	void counts(int* out) override {
		out[0] = pbc;
		out[1] = nbc;
		out[2] = ebc;
	}
	int keys(int* out) override {
		check_cuda_errors(cudaMemcpy(out, parts[rollid].active_keys, sizeof(int) * 3 * ebc, cudaMemcpyDeviceToHost));
		return ebc;
	}
	int grid(float* out) override {
		check_cuda_errors(cudaMemcpy(out, grids[0].handle.ptr, sizeof(float) * 256 * (size_t) nbc, cudaMemcpyDeviceToHost));
		return nbc;
	}
	int state(int m, float* out) override {
		const int R = rollid, Rn = R ^ 1;
		float* d_out = nullptr;
		int* d_cnt = nullptr;
		check_cuda_errors(cudaMalloc(&d_out, sizeof(float) * NCH * (size_t) counts_[m]));
		check_cuda_errors(cudaMalloc(&d_cnt, sizeof(int)));
		check_cuda_errors(cudaMemset(d_cnt, 0, sizeof(int)));
		export_state<<<pbc, 128, 0, st>>>(NCH, parts[R], parts[Rn], bins[R][m], bins[Rn][m], d_out, d_cnt);
		int n = 0;
		check_cuda_errors(cudaMemcpyAsync(&n, d_cnt, sizeof(int), cudaMemcpyDefault, st));
		check_cuda_errors(cudaStreamSynchronize(st));
		check_cuda_errors(cudaMemcpy(out, d_out, sizeof(float) * NCH * (size_t) n, cudaMemcpyDeviceToHost));
		cudaFree(d_out);
		cudaFree(d_cnt);
		return n;
	}
	float dt_now() override { return dt; }
};
}  // namespace

extern "C" {
#define REF_API __attribute__((visibility("default")))
REF_API int refgpu_domain_bits() { return config::DOMAIN_BITS; }
// route the kernels selected by `mask` (see g_swap) through libclaymore_b200 (path of the .so); mask 0 restores the reference's own
REF_API int refgpu_swap(const char* libpath, int mask) {
	if(mask == 0) {
		g_swap = 0;
		return 0;
	}
	void* h = dlopen(libpath, RTLD_NOW | RTLD_GLOBAL);
	if(!h) return -1;
#define X(n)                                       \
	g_b200_##n = (decltype(&::n)) dlsym(h, #n); \
	if(!g_b200_##n) return -2;
	B200_FUNCS(X)
#undef X
	g_swap = mask;
	return 0;
}
REF_API int refgpu_max_blocks() { return (int) config::G_MAX_ACTIVE_BLOCK; }
REF_API void* refgpu_create(int material, float dt_default) {
	switch(material) {
		case 0: return new Sim<MaterialE::J_FLUID>(dt_default);
		case 1: return new Sim<MaterialE::FIXED_COROTATED>(dt_default);
		case 2: return new Sim<MaterialE::SAND>(dt_default);
		default: return nullptr;
	}
}
REF_API int refgpu_init_model(void* h, const float* pos, int n, const float* v0, const float* params) { return static_cast<SimBase*>(h)->init_model(pos, n, v0, params); }
REF_API int refgpu_setup(void* h) { return static_cast<SimBase*>(h)->setup(); }
REF_API int refgpu_step(void* h, int n) { return static_cast<SimBase*>(h)->step(n); }
// wall-clock milliseconds of n sub-steps including the reference's host round trips (what a user of the reference waits for)
REF_API double refgpu_time_steps(void* h, int n) {
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0);
	const int err = static_cast<SimBase*>(h)->step(n);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return err ? -1.0 : (double) ms;
}
REF_API void refgpu_counts(void* h, int* out3) { static_cast<SimBase*>(h)->counts(out3); }
REF_API int refgpu_keys(void* h, int* out) { return static_cast<SimBase*>(h)->keys(out); }
REF_API int refgpu_grid(void* h, float* out) { return static_cast<SimBase*>(h)->grid(out); }
REF_API int refgpu_state(void* h, int model, float* out) { return static_cast<SimBase*>(h)->state(model, out); }
REF_API float refgpu_dt(void* h) { return static_cast<SimBase*>(h)->dt_now(); }
REF_API void refgpu_destroy(void* h) { delete static_cast<SimBase*>(h); }
}
