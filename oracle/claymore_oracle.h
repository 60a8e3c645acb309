/*
 * claymore_oracle.h -- CPU restatement of the claymore GMPM hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for claymore_b200.  It restates, in plain C on host memory, the
 * algorithm of the reference's fused G2P2G transfer and sparse-grid partition/update
 * (penn-graphics-research/claymore, Projects/GMPM).  Every function cites the reference
 * file:line it follows.  It operates on the SAME byte layouts as the reference containers
 * (AoSoA particle bins, 4^3 x 4-channel grid blocks, dense index table + active key list).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (claymore_b200/) never links, imports or calls it.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors.
 * The oracle is pinned against outputs of the reference's own kernels compiled unmodified
 * for sm_100a (oracle/_ref, built by oracle/build_ref.sh, run on the GPU box by
 * tests/golden/make_ref_golden.py); see tests/golden/README.md.
 */
#ifndef CLAYMORE_ORACLE_H
#define CLAYMORE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* MaterialE, Projects/GMPM/settings.h:23-29 */
enum { ORC_J_FLUID = 0, ORC_FIXED_COROTATED = 1, ORC_SAND = 2, ORC_NACC = 3 };

/* runtime form of namespace config, Projects/GMPM/settings.h:33-96 */
typedef struct orc_config {
	int domain_bits;   /* DOMAIN_BITS */
	int max_ppc;       /* G_MAX_PARTICLES_IN_CELL (power of two) */
	int boundary;      /* floor(G_BOUNDARY_CONDITION) */
	float gravity;     /* G_GRAVITY */
	float cfl;         /* CFL */
} orc_config;

/* ParticleBuffer<M> by-value fields, Projects/GMPM/particle_buffer.cuh:38-264 */
typedef struct orc_particle_buffer {
	int material;
	float* bins;                /* handle.ptr: bins of 32 particles, SoA, 512 B (fluid) / 2048 B */
	int* cell_particle_counts;  /* [blocks*64] */
	int* particle_bucket_sizes; /* [blocks] */
	int* cellbuckets;           /* [blocks*64*max_ppc] */
	int* blockbuckets;          /* [blocks*64*max_ppc] */
	int* bin_offsets;           /* [blocks] */
	float rho, volume, mass;
	/* J_FLUID */
	float bulk, gamma, viscosity;
	/* FIXED_COROTATED / SAND / NACC */
	float lambda, mu;
	/* SAND */
	float cohesion, beta, yield_surface;
	int volume_correction;
	/* NACC (beta shared) */
	float bm, xi, msqr;
	int hardening_on;
} orc_particle_buffer;

/* Partition<1>, Projects/GMPM/hash_table.cuh:75-135 */
typedef struct orc_partition {
	int* count;       /* 1 int */
	int* index_table; /* [G^3], sentinel -1 */
	int* active_keys; /* ivec3[capacity] */
	/* HaloPartition<1>, hash_table.cuh:27-73 */
	int* halo_count;
	char* halo_marks;
	int* overlap_marks;
	int* halo_blocks; /* ivec3[] */
} orc_partition;

void orc_set_num_threads(int n);
int orc_get_num_threads(void);

/* --- math --- */
void orc_svd3(const float* F /*col-major 9*/, float* U, float* S, float* V);
void orc_compute_stress(int material, const orc_particle_buffer* pb, float* F /*in/out*/, float* PF, float* log_jp);
void orc_bspline_weight(const orc_config* cfg, float p, float* w3);

/* --- init-only kernels --- */
void orc_activate_blocks(const orc_config* cfg, int n, const float* pos /*AoS xyz*/, orc_partition part);
void orc_build_particle_cell_buckets(const orc_config* cfg, int n, const float* pos, orc_particle_buffer pb, orc_partition part);
void orc_array_to_buffer(const orc_config* cfg, int block_count, const float* pos, orc_particle_buffer pb);
void orc_rasterize(const orc_config* cfg, int n, const float* pos, float* grid, orc_partition part, float mass, const float* v0);
void orc_init_adv_bucket(const orc_config* cfg, int block_count, const int* particle_bucket_sizes, int* buckets);

/* --- per-step kernels --- */
void orc_cell_bucket_to_block(const orc_config* cfg, int block_count, const int* cell_particle_counts, const int* cellbuckets, int* particle_bucket_sizes, int* buckets);
void orc_compute_bin_capacity(int block_count, const int* particle_bucket_sizes, int* bin_sizes);
void orc_exclusive_scan(int count, const int* in, int* out);
void orc_exclusive_scan_inverse(int num, const int* map, int* map_inv);
void orc_clear_grid(int block_count, float* grid);
void orc_register_neighbor_blocks(const orc_config* cfg, int block_count, orc_partition part);
void orc_register_exterior_blocks(const orc_config* cfg, int block_count, orc_partition part);
void orc_update_grid_velocity_query_max(const orc_config* cfg, int block_count, float* grid, orc_partition part, float dt, float* max_vel);
void orc_g2p2g(const orc_config* cfg, float dt, float new_dt, int block_count, orc_particle_buffer cur, orc_particle_buffer next, orc_partition prev_part, orc_partition part, const float* grid, float* next_grid);
void orc_mark_active_grid_blocks(int block_count, const float* grid, int* marks);
void orc_mark_active_particle_blocks(int block_count, const int* particle_bucket_sizes, int* marks);
void orc_update_partition(const orc_config* cfg, int block_count, const int* source_nos, orc_partition part, orc_partition next_part);
void orc_update_buckets(const orc_config* cfg, int block_count, const int* source_nos, orc_particle_buffer pb, orc_particle_buffer next_pb);
void orc_copy_selected_grid_blocks(const orc_config* cfg, int prev_block_count, const int* prev_blockids, orc_partition part, const int* marks, const float* prev_grid, float* grid);
int orc_retrieve_particle_buffer(const orc_config* cfg, int block_count, orc_partition part, orc_partition prev_part, orc_particle_buffer pb, orc_particle_buffer next_pb, float* out_pos);
void orc_reset_table(const orc_config* cfg, orc_partition part);

/* --- MGSP halo protocol (Projects/MGSP/halo_kernels.cuh:22-97) --- */
void orc_mark_overlapping_blocks(const orc_config* cfg, int block_count, int otherdid, const int* incoming_block_ids, orc_partition part, int* count, int* out_blockids);
void orc_collect_blockids_for_halo_reduction(const orc_config* cfg, int particle_block_count, orc_partition part);
void orc_collect_grid_blocks(const orc_config* cfg, int count, const int* blockids, const float* grid, orc_partition part, float* halo_grid);
void orc_reduce_grid_blocks(const orc_config* cfg, int count, const int* blockids, float* grid, orc_partition part, const float* halo_grid);

/* --- whole-pipeline driver (GmpmSimulator restated, Projects/GMPM/gmpm_simulator.cuh) --- */
typedef struct orc_sim orc_sim;
orc_sim* orc_sim_create(const orc_config* cfg, float dt_default, int max_blocks);
void orc_sim_destroy(orc_sim* s);
/* returns model id; params layout depends on material (see .c) ; params may be NULL for defaults */
int orc_sim_init_model(orc_sim* s, int material, const float* pos, int n, const float* v0);
void orc_sim_update_fr_parameters(orc_sim* s, int model, float rho, float vol, float ym, float pr);
void orc_sim_update_j_fluid_parameters(orc_sim* s, int model, float rho, float vol, float bulk, float gamma, float visc);
void orc_sim_update_nacc_parameters(orc_sim* s, int model, float rho, float vol, float ym, float pr, float beta, float xi);
void orc_sim_update_sand_parameters(orc_sim* s, int model, float rho, float vol, float ym, float pr);
void orc_sim_initial_setup(orc_sim* s);
/* one sub-step with fixed frame horizon `time_left` (pass a large value for "never binds") */
void orc_sim_step(orc_sim* s, float time_left);
int orc_sim_counts(orc_sim* s, int* pbc, int* nbc, int* ebc);
float orc_sim_dt(orc_sim* s);
float orc_sim_max_vel(orc_sim* s);
int orc_sim_retrieve(orc_sim* s, int model, float* out_pos);
/* access for parity tests */
const int* orc_sim_active_keys(orc_sim* s);         /* current partition keys, ivec3[ebc] */
const float* orc_sim_grid(orc_sim* s);              /* grid_blocks[0], nbc blocks valid */
int orc_sim_particle_state(orc_sim* s, int model, float* out /* n x channels, in bucket order */);

void orc_sim_get_buffer(orc_sim* s, int model, int which, orc_particle_buffer* out);
void orc_sim_get_partition(orc_sim* s, int which, orc_partition* out);
float* orc_sim_get_grid(orc_sim* s, int which);
long orc_sim_bin_capacity(orc_sim* s, int model);

#ifdef __cplusplus
}
#endif
#endif
