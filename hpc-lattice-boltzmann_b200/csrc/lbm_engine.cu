// lbm_engine.cu -- host side of the C ABI in include/lbm_b200.h: owns the device memory, streams,
// CUDA graphs and the slab decomposition, and launches the kernels of lbm_kernels.cuh.
//
// What it replaces in the reference (d2q9-bgk.c): the t_ocl object bundle (:35-67), its creation
// (:642-780), upload (:159-201), the `for tt` loop with its buffer ping-pong (:203-234),
// timestep/accelerate_flow/comp_func host wrappers (:294-424) including the per-step clFinish +
// 4*nx*ny-byte read-back + serial host sum (:408-423), download (:237-272) and release (:803-809).
//
// Design notes
//  * A lattice is a ring of row slabs, one per GPU.  Slab storage has a ghost row below and above;
//    the step kernel stores boundary-row outputs straight into the neighbour's ghost rows (peer
//    pointers over NVLink; its own ghost rows when the ring has one member).  Steps on different
//    slabs are ordered only against their two neighbours (events), never against the host.
//  * Between API calls the resident state is always the reference's canonical post-step state.
//    Inside lbm_run the inflow acceleration of step t+1 is folded into the store epilogue of step t;
//    the first step of a run is preceded by a stand-alone accelerate kernel and the last step of a
//    run does not pre-accelerate.
//  * The average-velocity reduction never leaves the device during a run: per-block double sums per
//    step, reduced per chunk of steps by a second kernel into a per-step totals array that is
//    copied back once, after the last step.
//  * On one GPU whole chunks of steps are replayed from a CUDA graph to take the launch overhead
//    out of small grids.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"
#include "lbm_kernels.cuh"

namespace {

thread_local std::string g_error;

int fail(const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return 1;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail("CUDA error during '%s' at %s:%d: %s", #call, __FILE__, __LINE__,           \
                  cudaGetErrorString(e_));                                                    \
  } while (0)

int env_int(const char* name, int dflt)
{
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

struct Slab {
  int        device = 0;
  int        rank = 0;        // position in the ring
  int        y0 = 0, rows = 0;
  long long  ps = 0;          // plane stride (floats)
  float*     buf[2] = {nullptr, nullptr};
  uint8_t*   flags = nullptr;
  double*    partials = nullptr;   // [chunk][nblk]
  double*    totals = nullptr;     // [totals_cap] per-step speed totals of this slab
  long long* counter = nullptr;
  long long  totals_cap = 0;
  int        nblk = 0;
  long long  nvec = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev_begin = nullptr, ev_end = nullptr;
  cudaEvent_t  ev_step[2] = {nullptr, nullptr};
  float*     ghost_lo[2][3] = {};  // [buffer][plane 4,7,8] destination of my first row
  float*     ghost_hi[2][3] = {};  // [buffer][plane 2,5,6] destination of my last row
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  bool       owns_accel_row = false;
  int        accel_row = 0;        // storage row of global row ny-2
};

}  // namespace

struct lbm_lattice {
  lbm_params p{};
  float a1 = 0, a2 = 0;
  long long tot_cells = 0;
  int world = 1;               // slabs in the ring
  std::vector<Slab> slabs;     // slabs driven by this process
  int cur = 0;                 // buffer holding the current state
  int host_y0 = 0;             // first lattice row of the caller's host planes (rank mode: the slab's)
  int vec = 4, tpb = 128, chunk = 128;
  bool use_graph = true;
  double last_ms = 0;
  long long last_launches = 0;
  std::string config;
};

namespace {

using lbm::StepArgs;

template <int VEC, int TPB>
cudaError_t launch_step_t(const StepArgs& a, int nblk, cudaStream_t st)
{
  lbm::lbm_step_kernel<VEC, TPB><<<nblk, TPB, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_step(int vec, int tpb, const StepArgs& a, int nblk, cudaStream_t st)
{
#define LBM_CASE(V, T) if (vec == V && tpb == T) return launch_step_t<V, T>(a, nblk, st);
  LBM_CASE(4, 64) LBM_CASE(4, 128) LBM_CASE(4, 256) LBM_CASE(4, 512)
  LBM_CASE(2, 64) LBM_CASE(2, 128) LBM_CASE(2, 256) LBM_CASE(2, 512)
  LBM_CASE(1, 64) LBM_CASE(1, 128) LBM_CASE(1, 256) LBM_CASE(1, 512)
#undef LBM_CASE
  return cudaErrorInvalidValue;
}

StepArgs make_args(const lbm_lattice* h, const Slab& s, int cur, int fuse, int slot)
{
  StepArgs a{};
  a.src = s.buf[cur];
  a.dst = s.buf[cur ^ 1];
  a.flags = s.flags;
  a.ps = s.ps;
  a.nvec = s.nvec;
  a.nx = h->p.nx;
  a.rows = s.rows;
  a.nxv = h->p.nx / h->vec;
  a.omega = h->p.omega;
  a.a1 = h->a1;
  a.a2 = h->a2;
  a.fuse_accel = fuse;
  for (int i = 0; i < 3; i++) {
    a.ghost_lo[i] = s.ghost_lo[cur ^ 1][i];
    a.ghost_hi[i] = s.ghost_hi[cur ^ 1][i];
  }
  a.partials = s.partials + (long long)slot * s.nblk;
  return a;
}

int launch_accelerate(lbm_lattice* h, Slab& s, int cur)
{
  if (!s.owns_accel_row) return 0;
  const int nx = h->p.nx;
  lbm::accelerate_row_kernel<<<(nx + 255) / 256, 256, 0, s.stream>>>(
      s.buf[cur], s.flags, s.ps, nx, s.accel_row, s.rows, h->a1, h->a2, s.ghost_lo[cur][1],
      s.ghost_lo[cur][2], s.ghost_hi[cur][1], s.ghost_hi[cur][2]);
  CK(cudaGetLastError());
  h->last_launches++;
  return 0;
}

int launch_halo_push(lbm_lattice* h, Slab& s, int cur)
{
  const int nx = h->p.nx;
  lbm::halo_push_kernel<<<(nx + 255) / 256, 256, 0, s.stream>>>(
      s.buf[cur], s.ps, nx, s.rows, s.ghost_lo[cur][0], s.ghost_lo[cur][1], s.ghost_lo[cur][2],
      s.ghost_hi[cur][0], s.ghost_hi[cur][1], s.ghost_hi[cur][2]);
  CK(cudaGetLastError());
  return 0;
}

int sync_all(lbm_lattice* h)
{
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    CK(cudaStreamSynchronize(s.stream));
  }
  return 0;
}

// after the resident state changed from outside (upload / init): fill every ghost row
int refresh_ghosts(lbm_lattice* h)
{
  if (sync_all(h)) return 1;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    if (launch_halo_push(h, s, h->cur)) return 1;
  }
  return sync_all(h);
}

int ensure_totals(lbm_lattice* h, long long iters)
{
  for (auto& s : h->slabs) {
    if (s.totals_cap >= iters) continue;
    CK(cudaSetDevice(s.device));
    if (s.totals) CK(cudaFree(s.totals));
    s.totals = nullptr;
    const long long cap = std::max<long long>(iters, 1024);
    CK(cudaMalloc(&s.totals, sizeof(double) * cap));
    s.totals_cap = cap;
  }
  return 0;
}

// one chunk of `h->chunk` pre-accelerating steps + its reduction, captured once per buffer parity
int build_graph(lbm_lattice* h, Slab& s, int cur)
{
  cudaGraph_t g = nullptr;
  CK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
  int c = cur;
  for (int i = 0; i < h->chunk; i++) {
    const StepArgs a = make_args(h, s, c, 1, i);
    CK(launch_step(h->vec, h->tpb, a, s.nblk, s.stream));
    c ^= 1;
  }
  lbm::reduce_partials_kernel<<<h->chunk, 256, 0, s.stream>>>(s.partials, s.nblk, s.totals, s.counter);
  CK(cudaGetLastError());
  lbm::advance_counter_kernel<<<1, 1, 0, s.stream>>>(s.counter, h->chunk);
  CK(cudaGetLastError());
  CK(cudaStreamEndCapture(s.stream, &g));
  CK(cudaGraphInstantiate(&s.graph[cur], g, 0));
  CK(cudaGraphDestroy(g));
  return 0;
}

void drop_graphs(Slab& s)
{
  for (int i = 0; i < 2; i++)
    if (s.graph[i]) { cudaGraphExecDestroy(s.graph[i]); s.graph[i] = nullptr; }
}

int run_impl(lbm_lattice* h, int iters, double* av_out)
{
  h->last_ms = 0;
  h->last_launches = 0;
  if (iters < 0) return fail("lbm_run: negative iteration count");
  if (iters == 0) return 0;
  const size_t nslab = h->slabs.size();
  const long long totals_before = h->slabs[0].totals_cap;
  if (ensure_totals(h, iters)) return 1;
  if (h->slabs[0].totals_cap != totals_before)
    for (auto& s : h->slabs) drop_graphs(s);   // the captured reduce node holds the old pointer

  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    CK(cudaMemsetAsync(s.counter, 0, sizeof(long long), s.stream));
    CK(cudaEventRecord(s.ev_begin, s.stream));
    if (launch_accelerate(h, s, h->cur)) return 1;
  }

  int remaining = iters;
  int cur = h->cur;
  if (nslab == 1 && h->use_graph) {
    Slab& s = h->slabs[0];
    while (remaining > h->chunk) {
      if (!s.graph[cur] && build_graph(h, s, cur)) return 1;
      CK(cudaGraphLaunch(s.graph[cur], s.stream));
      h->last_launches += h->chunk + 2;
      remaining -= h->chunk;                    // chunk is even: parity unchanged
    }
  }

  // remaining steps as plain launches, reduced chunk by chunk; the very last step of the run
  // leaves the state un-accelerated
  long long step_no = 0;
  while (remaining > 0) {
    const int n = std::min(remaining, h->chunk);
    for (int i = 0; i < n; i++, step_no++) {
      const int fuse = (remaining - i) > 1;
      for (size_t k = 0; k < nslab; k++) {
        Slab& s = h->slabs[k];
        CK(cudaSetDevice(s.device));
        if (nslab > 1) {
          // my neighbours must have finished the previous step: their stores into my ghost rows
          // are complete and they no longer read the ghost rows I am about to overwrite
          if (step_no > 0) {
            const Slab& lo = h->slabs[(k + nslab - 1) % nslab];
            const Slab& hi = h->slabs[(k + 1) % nslab];
            CK(cudaStreamWaitEvent(s.stream, lo.ev_step[(step_no - 1) & 1], 0));
            CK(cudaStreamWaitEvent(s.stream, hi.ev_step[(step_no - 1) & 1], 0));
          }
        }
        const StepArgs a = make_args(h, s, cur, fuse, i);
        CK(launch_step(h->vec, h->tpb, a, s.nblk, s.stream));
        if (nslab > 1) CK(cudaEventRecord(s.ev_step[step_no & 1], s.stream));
      }
      h->last_launches += (long long)nslab;
      cur ^= 1;
    }
    for (auto& s : h->slabs) {
      CK(cudaSetDevice(s.device));
      lbm::reduce_partials_kernel<<<n, 256, 0, s.stream>>>(s.partials, s.nblk, s.totals, s.counter);
      CK(cudaGetLastError());
      lbm::advance_counter_kernel<<<1, 1, 0, s.stream>>>(s.counter, n);
      CK(cudaGetLastError());
    }
    h->last_launches += 2 * (long long)nslab;
    remaining -= n;
  }
  h->cur = cur;

  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    CK(cudaEventRecord(s.ev_end, s.stream));
  }
  if (sync_all(h)) return 1;
  float ms_max = 0;
  for (auto& s : h->slabs) {
    float ms = 0;
    CK(cudaSetDevice(s.device));
    CK(cudaEventElapsedTime(&ms, s.ev_begin, s.ev_end));
    ms_max = std::max(ms_max, ms);
  }
  h->last_ms = ms_max;

  if (av_out) {
    std::vector<double> tmp((size_t)iters);
    std::fill(av_out, av_out + iters, 0.0);
    for (auto& s : h->slabs) {          // fixed slab order: the cross-GPU sum is deterministic
      CK(cudaSetDevice(s.device));
      CK(cudaMemcpy(tmp.data(), s.totals, sizeof(double) * iters, cudaMemcpyDeviceToHost));
      for (int t = 0; t < iters; t++) av_out[t] += tmp[t];
    }
    const double inv = (double)h->tot_cells;
    for (int t = 0; t < iters; t++) av_out[t] /= inv;
  }
  return 0;
}

int create_impl(lbm_lattice** out, const lbm_params* p, const int* obstacles, int first_device,
                int nslab)
{
  if (!out || !p || !obstacles) return fail("lbm_create: null argument");
  *out = nullptr;
  if (p->nx < 1 || p->ny < 1) return fail("lbm_create: bad grid %d x %d", p->nx, p->ny);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail("lbm_create: no CUDA device available (this library has no CPU fallback)");
  if (nslab < 1 || first_device < 0 || first_device + nslab > ndev)
    return fail("lbm_create: %d GPU(s) requested from device %d but %d visible", nslab, first_device, ndev);
  if (nslab > 1 && p->ny / nslab < 3)
    return fail("lbm_create: %d rows cannot be split into %d slabs of >= 3 rows", p->ny, nslab);

  lbm_lattice* h = new lbm_lattice();
  h->p = *p;
  // kernels.cl:17-18: float product, double divide, rounded to float
  h->a1 = (float)((double)(p->density * p->accel) / 9.0);
  h->a2 = (float)((double)(p->density * p->accel) / 36.0);
  h->world = nslab;
  const int nx = p->nx, ny = p->ny;

  int vec = (nx % 4 == 0) ? 4 : (nx % 2 == 0) ? 2 : 1;
  const int want_vec = env_int("LBM_VEC", vec);
  if ((want_vec == 1 || want_vec == 2 || want_vec == 4) && nx % want_vec == 0) vec = want_vec;
  h->vec = vec;
  // 128-thread blocks measured best on B200 (profiles/r1_sweep.md): 7 resident blocks per SM at
  // 70 registers and a finer-grained tail than 256/512
  const int want_tpb = env_int("LBM_TPB", 128);
  h->tpb = (want_tpb == 64 || want_tpb == 256 || want_tpb == 512) ? want_tpb : 128;
  h->chunk = std::max(2, env_int("LBM_CHUNK", 128)) & ~1;
  h->use_graph = env_int("LBM_GRAPH", 1) != 0;
  const int pad = std::max(0, env_int("LBM_PLANE_PAD", 0));

  h->slabs.resize(nslab);
  auto bail = [&](int) { lbm_destroy(h); return 1; };

  for (int k = 0; k < nslab; k++) {
    Slab& s = h->slabs[k];
    s.device = first_device + k;
    s.rank = k;
    lbm_slab_rows(ny, nslab, k, &s.y0, &s.rows);
    const long long cells = (long long)(s.rows + 2) * nx;
    s.ps = ((cells + 31) / 32) * 32 + ((long long)pad / 32) * 32;
    s.nvec = (long long)s.rows * (nx / vec);
    if (s.nvec >= (1LL << 31)) { fail("lbm_create: slab too large for 32-bit work index"); return bail(0); }
    s.nblk = (int)((s.nvec + h->tpb - 1) / h->tpb);
    s.owns_accel_row = (ny - 2 >= s.y0 && ny - 2 < s.y0 + s.rows);
    s.accel_row = ny - 2 - s.y0 + 1;

    if (cudaSetDevice(s.device) != cudaSuccess) { fail("cudaSetDevice(%d) failed", s.device); return bail(0); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, s.device) != cudaSuccess || prop.major < 10) {
      fail("device %d is not an sm_100-class GPU (this library is built for sm_100a only)", s.device);
      return bail(0);
    }
#define CKB(call)                                                                             \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      fail("CUDA error during '%s' at %s:%d: %s", #call, __FILE__, __LINE__,                  \
           cudaGetErrorString(e_));                                                           \
      return bail(0);                                                                         \
    }                                                                                         \
  } while (0)
    CKB(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CKB(cudaEventCreate(&s.ev_begin));
    CKB(cudaEventCreate(&s.ev_end));
    CKB(cudaEventCreateWithFlags(&s.ev_step[0], cudaEventDisableTiming));
    CKB(cudaEventCreateWithFlags(&s.ev_step[1], cudaEventDisableTiming));
    for (int b = 0; b < 2; b++) {
      CKB(cudaMalloc(&s.buf[b], sizeof(float) * 9 * s.ps));
      CKB(cudaMemset(s.buf[b], 0, sizeof(float) * 9 * s.ps));
    }
    CKB(cudaMalloc(&s.flags, (size_t)cells));
    CKB(cudaMalloc(&s.partials, sizeof(double) * (size_t)h->chunk * s.nblk));
    CKB(cudaMalloc(&s.counter, sizeof(long long)));
    CKB(cudaMemset(s.counter, 0, sizeof(long long)));

    // flags: bit 0 obstacle, bit 1 fluid cell of the accelerated row (global ny-2)
    std::vector<uint8_t> fl((size_t)cells, 0);
    for (int r = 0; r < s.rows; r++) {
      const int gy = s.y0 + r;
      const int* orow = obstacles + (size_t)gy * nx;
      uint8_t* frow = fl.data() + (size_t)(r + 1) * nx;
      for (int x = 0; x < nx; x++) {
        const bool ob = orow[x] != 0;
        frow[x] = (uint8_t)((ob ? lbm::FLAG_OBSTACLE : 0) | ((!ob && gy == ny - 2) ? lbm::FLAG_ACCEL : 0));
        h->tot_cells += !ob;
      }
    }
    CKB(cudaMemcpy(s.flags, fl.data(), (size_t)cells, cudaMemcpyHostToDevice));
  }

  // ring wiring: where each slab's boundary rows land
  for (int k = 0; k < nslab; k++) {
    Slab& s = h->slabs[k];
    Slab& lo = h->slabs[(k + nslab - 1) % nslab];
    Slab& hi = h->slabs[(k + 1) % nslab];
    static const int lo_planes[3] = {4, 7, 8}, hi_planes[3] = {2, 5, 6};
    for (int b = 0; b < 2; b++)
      for (int i = 0; i < 3; i++) {
        s.ghost_lo[b][i] = lo.buf[b] + lo_planes[i] * lo.ps + (long long)(lo.rows + 1) * nx;
        s.ghost_hi[b][i] = hi.buf[b] + hi_planes[i] * hi.ps;
      }
    if (nslab > 1) {
      CKB(cudaSetDevice(s.device));
      for (const Slab* nb : {&lo, &hi}) {
        if (nb->device == s.device) continue;
        int can = 0;
        CKB(cudaDeviceCanAccessPeer(&can, s.device, nb->device));
        if (!can) { fail("GPU %d cannot access GPU %d peer memory", s.device, nb->device); return bail(0); }
        cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) { fail("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return bail(0); }
      }
    }
  }
#undef CKB

  char cfg[256];
  snprintf(cfg, sizeof cfg, "vec=%d tpb=%d chunk=%d graph=%d slabs=%d plane_stride=%lld",
           h->vec, h->tpb, h->chunk, (int)h->use_graph, nslab, h->slabs[0].ps);
  h->config = cfg;
  *out = h;
  return 0;
}

}  // namespace

extern "C" {

const char* lbm_last_error(void) { return g_error.c_str(); }

int lbm_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int lbm_slab_rows(int ny, int world, int rank, int* y0, int* rows)
{
  if (world < 1 || rank < 0 || rank >= world || ny < world) return fail("lbm_slab_rows: bad arguments");
  const int base = ny / world, extra = ny % world;       // the first `extra` slabs get one more row
  if (y0) *y0 = rank * base + std::min(rank, extra);
  if (rows) *rows = base + (rank < extra ? 1 : 0);
  return 0;
}

int lbm_create(lbm_lattice** out, const lbm_params* params, const int* obstacles, int ngpus)
{
  return create_impl(out, params, obstacles, env_int("LBM_FIRST_DEVICE", 0), ngpus);
}

int lbm_create_rank(lbm_lattice** out, const lbm_params* params, const int* obstacles_slab,
                    int rank, int world, int device, const void* nccl_unique_id)
{
  (void)nccl_unique_id;
  if (world == 1 && rank == 0) return create_impl(out, params, obstacles_slab, device, 1);
  return fail("lbm_create_rank: world > 1 not available in this build");
}

int lbm_comm_unique_id(void* out128)
{
  (void)out128;
  return fail("lbm_comm_unique_id: not available in this build");
}

void lbm_destroy(lbm_lattice* h)
{
  if (!h) return;
  for (auto& s : h->slabs) {
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
    drop_graphs(s);
    for (int b = 0; b < 2; b++) if (s.buf[b]) cudaFree(s.buf[b]);
    if (s.flags) cudaFree(s.flags);
    if (s.partials) cudaFree(s.partials);
    if (s.totals) cudaFree(s.totals);
    if (s.counter) cudaFree(s.counter);
    if (s.ev_begin) cudaEventDestroy(s.ev_begin);
    if (s.ev_end) cudaEventDestroy(s.ev_end);
    for (int i = 0; i < 2; i++) if (s.ev_step[i]) cudaEventDestroy(s.ev_step[i]);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  delete h;
}

int lbm_init_equilibrium(lbm_lattice* h)
{
  if (!h) return fail("null handle");
  // d2q9-bgk.c:573-575: float <- density (float) * 4.0 / 9.0 evaluated in double
  const float w0 = (float)((double)h->p.density * 4.0 / 9.0);
  const float w1 = (float)((double)h->p.density / 9.0);
  const float w2 = (float)((double)h->p.density / 36.0);
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    const long long cells = (long long)(s.rows + 2) * h->p.nx;
    lbm::init_equilibrium_kernel<<<148 * 8, 256, 0, s.stream>>>(s.buf[h->cur], s.ps, cells, w0, w1, w2);
    CK(cudaGetLastError());
  }
  return sync_all(h);
}

int lbm_upload(lbm_lattice* h, const float* const cells[9])
{
  if (!h || !cells) return fail("null argument");
  const int nx = h->p.nx;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    for (int k = 0; k < 9; k++)
      CK(cudaMemcpyAsync(s.buf[h->cur] + k * s.ps + nx, cells[k] + (size_t)(s.y0 - h->host_y0) * nx,
                         sizeof(float) * (size_t)s.rows * nx, cudaMemcpyHostToDevice, s.stream));
  }
  return refresh_ghosts(h);
}

int lbm_download(lbm_lattice* h, float* const cells[9])
{
  if (!h || !cells) return fail("null argument");
  const int nx = h->p.nx;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    for (int k = 0; k < 9; k++)
      CK(cudaMemcpyAsync(cells[k] + (size_t)(s.y0 - h->host_y0) * nx, s.buf[h->cur] + k * s.ps + nx,
                         sizeof(float) * (size_t)s.rows * nx, cudaMemcpyDeviceToHost, s.stream));
  }
  return sync_all(h);
}

int lbm_run_f64(lbm_lattice* h, int iters, double* av_vels)
{
  if (!h) return fail("null handle");
  if (av_vels) return run_impl(h, iters, av_vels);
  std::vector<double> scratch((size_t)std::max(iters, 1));
  return run_impl(h, iters, scratch.data());
}

int lbm_run(lbm_lattice* h, int iters, float* av_vels)
{
  if (!h) return fail("null handle");
  std::vector<double> tmp((size_t)std::max(iters, 1));
  if (run_impl(h, iters, tmp.data())) return 1;
  if (av_vels) for (int t = 0; t < iters; t++) av_vels[t] = (float)tmp[t];
  return 0;
}

int lbm_step(lbm_lattice* h, float* av_vel) { return lbm_run(h, 1, av_vel); }

int lbm_av_velocity(lbm_lattice* h, float* av_vel)
{
  if (!h || !av_vel) return fail("null argument");
  double total = 0;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    const int nblk = std::min(s.nblk, 148 * 8);
    lbm::av_velocity_kernel<256><<<nblk, 256, 0, s.stream>>>(s.buf[h->cur], s.flags, s.ps, h->p.nx,
                                                           s.rows, s.partials);
    CK(cudaGetLastError());
    std::vector<double> part((size_t)nblk);
    CK(cudaMemcpyAsync(part.data(), s.partials, sizeof(double) * nblk, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (double v : part) total += v;
  }
  *av_vel = (float)(total / (double)h->tot_cells);
  return 0;
}

int lbm_macroscopic(lbm_lattice* h, float* ux, float* uy, float* speed, float* pressure)
{
  if (!h || !ux || !uy || !speed || !pressure) return fail("null argument");
  const int nx = h->p.nx;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    const size_t n = (size_t)s.rows * nx;
    float* scratch = s.buf[h->cur ^ 1];          // the idle buffer: 4 planes of it are plenty
    lbm::macroscopic_kernel<<<148 * 8, 256, 0, s.stream>>>(s.buf[h->cur], s.flags, s.ps, nx, s.rows,
                                                         h->p.density, scratch, scratch + s.ps,
                                                         scratch + 2 * s.ps, scratch + 3 * s.ps);
    CK(cudaGetLastError());
    float* outs[4] = {ux, uy, speed, pressure};
    for (int k = 0; k < 4; k++)
      CK(cudaMemcpyAsync(outs[k] + (size_t)(s.y0 - h->host_y0) * nx, scratch + k * s.ps, sizeof(float) * n,
                         cudaMemcpyDeviceToHost, s.stream));
  }
  return sync_all(h);
}

double lbm_last_run_ms(const lbm_lattice* h) { return h ? h->last_ms : 0.0; }
long long lbm_last_run_launches(const lbm_lattice* h) { return h ? h->last_launches : 0; }
long long lbm_tot_cells(const lbm_lattice* h) { return h ? h->tot_cells : 0; }
const char* lbm_config_string(const lbm_lattice* h) { return h ? h->config.c_str() : ""; }

int lbm_local_slab(const lbm_lattice* h, int* y0, int* rows)
{
  if (!h || h->slabs.empty()) return fail("null handle");
  if (y0) *y0 = h->slabs.front().y0;
  int total = 0;
  for (auto& s : h->slabs) total += s.rows;
  if (rows) *rows = total;
  return 0;
}

int lbm_host_alloc(void** out, unsigned long long bytes)
{
  if (!out) return fail("null argument");
  CK(cudaMallocHost(out, (size_t)bytes));
  return 0;
}

void lbm_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
