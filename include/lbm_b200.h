/*
 * lbm_b200.h -- C ABI of the B200-native D2Q9-BGK lattice-Boltzmann timestep engine.
 *
 * This is the drop-in boundary for the per-timestep path of AlexDalt/HPC-Lattice-Boltzmann
 * (reference file d2q9-bgk.c).  The reference has no FFI layer; the seam it does have is the
 * bundle of OpenCL objects `t_ocl` (d2q9-bgk.c:35-67) that main() creates, feeds, steps and
 * reads back.  Each entry point below names the reference interface it replaces.  Plain C:
 * pointers and sizes only, no CUDA or torch types.  All device memory lives behind the opaque
 * handle; all host memory is caller-owned.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; lbm_last_error() then returns a
 *    message for the calling thread (the reference's print-and-exit convention, checkError
 *    d2q9-bgk.c:923-931 / die :933-939, is kept by the host program, not by the library);
 *  - populations are float32, SoA: `cells[k]` is plane k (speed k of kernels.cl:90-98:
 *    0 rest, 1 E, 2 N, 3 W, 4 S, 5 NE, 6 NW, 7 SW, 8 SE), row-major [y*nx + x], nx*ny floats;
 *  - obstacles are int32, non-zero = blocked, same indexing (d2q9-bgk.c:627);
 *  - single host thread per handle; calls are synchronous (they return when the GPU work is done),
 *    but no host synchronisation happens between the timesteps inside lbm_run;
 *  - there is NO CPU fallback: without a usable sm_100 device lbm_create* fails.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lbm_lattice lbm_lattice;

/* replaces t_param's physics fields (d2q9-bgk.c:23-32); maxIters / reynolds_dim stay with the caller */
typedef struct {
  int   nx;        /* cells in x */
  int   ny;        /* cells in y */
  float density;   /* density per link */
  float accel;     /* density redistribution */
  float omega;     /* relaxation parameter */
} lbm_params;

/*
 * Create an engine for the whole nx*ny grid on `ngpus` GPUs of this process (devices 0..ngpus-1,
 * row slabs, NVLink peer stores for the halos).  obstacles: nx*ny ints.
 * Replaces: selectOpenCLDevice + context/queue/program/kernels/20 clCreateBuffer
 * (d2q9-bgk.c:642-780) and the obstacle upload (:197-201); tot_cells count (:146-152).
 */
int lbm_create(lbm_lattice** out, const lbm_params* params, const int* obstacles, int ngpus);

/*
 * One-process-per-GPU flavour: this process owns row slab `rank` of `world` on CUDA device
 * `device`.  obstacles_slab: the slab's rows only (lbm_slab_rows tells which).  nccl_unique_id:
 * the 128 bytes of an ncclUniqueId made by rank 0 (lbm_comm_unique_id) and distributed by the
 * caller; NULL when world == 1.  Collective: all ranks must call it.
 */
int lbm_create_rank(lbm_lattice** out, const lbm_params* params, const int* obstacles_slab,
                    int rank, int world, int device, const void* nccl_unique_id);
int lbm_comm_unique_id(void* out128);

/* rows [y0, y0+rows) owned by slab `rank` of `world` for a grid of ny rows (pure arithmetic) */
int lbm_slab_rows(int ny, int world, int rank, int* y0, int* rows);

void lbm_destroy(lbm_lattice* h);   /* replaces finalise's cl releases, d2q9-bgk.c:803-809 */

/* initial equilibrium state generated on the device; replaces the host fill (:573-594) + upload */
int lbm_init_equilibrium(lbm_lattice* h);

/* replaces the 9 clEnqueueWriteBuffer of d2q9-bgk.c:159-195.  In rank mode the planes hold the
 * slab's rows only (rows*nx floats each). */
int lbm_upload(lbm_lattice* h, const float* const cells[9]);

/* replaces the 9 clEnqueueReadBuffer of d2q9-bgk.c:237-272 (and reads the NEWEST buffer) */
int lbm_download(lbm_lattice* h, float* const cells[9]);

/*
 * `float timestep(params, cells, tmp_cells, ocl, tot_cells)` (d2q9-bgk.c:294-298): one
 * accelerate_flow + propagate + rebound + collision + av_velocity; *av_vel = the step's average
 * speed over non-blocked cells.
 */
int lbm_step(lbm_lattice* h, float* av_vel);

/* the whole `for tt` loop (d2q9-bgk.c:206-234): iters timesteps with no host involvement,
 * av_vels[0..iters) written at the end.  av_vels may be NULL.  How the loop is mapped onto the GPU is
 * decided at lbm_create from the lattice alone (lbm_config_string tells): one persistent launch for
 * lattices that fit the SMs' shared memory, two timesteps per launch for slabs that stream from HBM, one
 * launch per timestep otherwise -- the results are bit-identical whichever runs. */
int lbm_run(lbm_lattice* h, int iters, float* av_vels);
int lbm_run_f64(lbm_lattice* h, int iters, double* av_vels);   /* same, un-narrowed averages */

/* `float av_velocity(params, cells, obstacles, ocl)` (d2q9-bgk.c:426-475) on the resident state */
int lbm_av_velocity(lbm_lattice* h, float* av_vel);

/* `float total_density(params, cells)` (d2q9-bgk.c:822-838): sum of every population of every cell of
 * the resident state, accumulated in double on the device (the reference's DEBUG conservation check,
 * which reads stale host data there; here it looks at the live state) */
int lbm_total_density(lbm_lattice* h, double* total);

/* final-state fields of write_values (d2q9-bgk.c:857-897) computed on the device: u_x, u_y, |u|,
 * pressure, nx*ny floats each (slab rows in rank mode); obstacle cells get 0,0,0,density/3 */
int lbm_macroscopic(lbm_lattice* h, float* ux, float* uy, float* speed, float* pressure);

/* introspection for benchmarks: CUDA-event time of the last lbm_run/lbm_step on its own stream,
 * kernels launched by it, number of non-blocked cells, the slab this handle owns */
double      lbm_last_run_ms(const lbm_lattice* h);
long long   lbm_last_run_launches(const lbm_lattice* h);
long long   lbm_tot_cells(const lbm_lattice* h);
int         lbm_local_slab(const lbm_lattice* h, int* y0, int* rows);
const char* lbm_config_string(const lbm_lattice* h);

/* measurement aid: GB/s (read + write) of a device-to-device copy of two `bytes`-sized buffers repeated
 * `reps` times inside one launch; with 2 x bytes well below the 126 MB L2 this is the L2 bandwidth that
 * bounds a lattice small enough to live there (the 1024 x 1024 case) */
int lbm_probe_l2_copy(unsigned long long bytes, int reps, double* gbs);

/* pinned host memory for callers that want full-speed transfers (optional) */
int  lbm_host_alloc(void** out, unsigned long long bytes);
void lbm_host_free(void* p);

const char* lbm_last_error(void);
int         lbm_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
