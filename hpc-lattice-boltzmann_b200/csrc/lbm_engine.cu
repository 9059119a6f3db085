// lbm_engine.cu -- host side of the C ABI in include/lbm_b200.h: owns the device memory, streams,
// CUDA graphs, the slab decomposition and the inter-GPU plumbing, and launches the kernels of
// lbm_kernels.cuh.
//
// What it replaces in the reference (d2q9-bgk.c): the t_ocl object bundle (:35-67), its creation
// (:642-780), upload (:159-201), the `for tt` loop with its buffer ping-pong (:203-234),
// timestep/accelerate_flow/comp_func host wrappers (:294-424) including the per-step clFinish +
// 4*nx*ny-byte read-back + serial host sum (:408-423), download (:237-272) and release (:803-809).
//
// Design notes
//  * A lattice is a ring of row slabs, one per GPU.  Slab storage has a ghost row below and above;
//    the step kernel stores boundary-row outputs straight into the neighbour's ghost rows: its own
//    ghost rows when the ring has one member, peer pointers over NVLink otherwise (direct peer
//    access when one process drives all GPUs, CUDA-IPC mappings when there is one process per GPU).
//    Steps on different slabs are ordered only against their two ring neighbours -- by stream
//    events inside one process, by release/acquire counters in peer memory across processes --
//    never against the host.  LBM_HALO=nccl switches the multi-process halo to ncclSend/ncclRecv.
//  * Between API calls the resident state is always the reference's canonical post-step state.
//    Inside lbm_run the inflow acceleration of step t+1 is folded into the store epilogue of step t;
//    the first step of a run is preceded by a stand-alone accelerate kernel and the last step of a
//    run does not pre-accelerate.
//  * The average-velocity reduction never leaves the device during a run: per-block double sums per
//    step, reduced per chunk of steps by a second kernel into a per-step totals array.  After the
//    last step the per-slab totals are combined in slab order (all-gathered first when the slabs
//    live in different processes), so the result does not depend on timing.
//  * On one GPU whole chunks of steps are replayed from a CUDA graph to take the launch overhead
//    out of small grids.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

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

// ---- NCCL, bound at run time ------------------------------------------------------------------
// dlopen by soname: inside a PyTorch process this resolves to the libnccl.so.2 torch already
// loaded, in the plain C program to the system one.  Only the one-process-per-GPU mode needs it.
struct NcclApi {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api()
{
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.so ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.so) break;
  }
  if (!api.so) return nullptr;
#define BIND(field, sym)                                                       \
  *(void**)(&api.field) = dlsym(api.so, sym);                                  \
  if (!api.field) { api.so = nullptr; return nullptr; }
  BIND(GetUniqueId, "ncclGetUniqueId") BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy") BIND(AllReduce, "ncclAllReduce")
  BIND(AllGather, "ncclAllGather") BIND(Send, "ncclSend") BIND(Recv, "ncclRecv")
  BIND(GroupStart, "ncclGroupStart") BIND(GroupEnd, "ncclGroupEnd")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  return &api;
}

#define NK(call)                                                                              \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess)                                                                    \
      return fail("NCCL error during '%s' at %s:%d: %s", #call, __FILE__, __LINE__,           \
                  nccl_api()->GetErrorString(r_));                                            \
  } while (0)

enum { HALO_P2P = 0, HALO_NCCL = 1 };

// one-process-per-GPU plumbing of a slab
struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
  int halo = HALO_P2P;
  void* peer_lo = nullptr;       // IPC mapping of the lower neighbour's slab allocation
  void* peer_hi = nullptr;       // ... upper neighbour's (same mapping when world == 2)
  unsigned* peer_lo_flag = nullptr;   // lower neighbour's "my upper neighbour has done N steps"
  unsigned* peer_hi_flag = nullptr;   // upper neighbour's "my lower neighbour has done N steps"
  unsigned steps_done = 0;       // steps completed since creation (all ranks advance together)
  float* dummy_ghost = nullptr;  // NCCL halo: the kernel's ghost stores go nowhere useful
  long long* scratch64 = nullptr;
  bool ready = false;            // fully attached: destroy may run its closing barrier
  bool in_kernel = false;        // ring ordering done by the step kernel's boundary blocks
  // LBM_REDUCE=step: one 8-byte ncclAllReduce per timestep (on a side stream, so the next step
  // does not wait for it) instead of one all-gather of the per-step totals after the run
  bool per_step_allreduce = false;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_side = nullptr;
  double* step_sums = nullptr;
  long long step_sums_cap = 0;
};

struct Slab {
  int        device = 0;
  int        rank = 0;        // position in the ring
  int        y0 = 0, rows = 0;
  long long  ps = 0;          // plane stride (floats)
  char*      base = nullptr;  // one allocation: buffer 0 | buffer 1 | sync words
  float*     buf[2] = {nullptr, nullptr};
  unsigned*  sync = nullptr;  // [0] steps done by my lower neighbour, [1] by my upper, [2] timeout
  float*     strip_lo = nullptr;   // two-step passes: t+1 rows [lower neighbour's last | my 1 | my 2]
  float*     strip_hi = nullptr;   //                           [my rows-1 | my rows | upper neighbour's first]
  long long  pse = 0;              // strip plane stride (floats)
  float*     nb_lo[3] = {};        // lower neighbour's strip_hi row 2, planes 4,7,8
  float*     nb_hi[3] = {};        // upper neighbour's strip_lo row 0, planes 2,5,6
  int        np = 0;               // partial-sum slots per step
  int        tiles_x = 0, tiles_y = 0;
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

const int LO_PLANES[3] = {4, 7, 8};   // pulled by the row below (kernels.cl:94,97,98)
const int HI_PLANES[3] = {2, 5, 6};   // pulled by the row above (kernels.cl:92,95,96)

long long plane_stride(int rows, int nx, int pad)
{
  const long long cells = (long long)(rows + 2) * nx;
  return ((cells + 31) / 32) * 32 + ((long long)pad / 32) * 32;
}

long long strip_stride(int nx) { return ((3LL * nx + 31) / 32) * 32; }

// buffer 0 | buffer 1 | 256 B of sync words | strip_lo | strip_hi
size_t slab_bytes(long long ps, int nx) { return sizeof(float) * (18 * (size_t)ps + 64 + 18 * (size_t)strip_stride(nx)); }

float* strip_of(char* base, long long ps, int nx, int hi)
{
  return reinterpret_cast<float*>(base) + 18 * ps + 64 + (hi ? 9 * strip_stride(nx) : 0);
}

}  // namespace

struct lbm_lattice {
  lbm_params p{};
  float a1 = 0, a2 = 0;
  long long tot_cells = 0;
  int world = 1;               // slabs in the ring
  std::vector<Slab> slabs;     // slabs driven by this process
  Comm* comm = nullptr;        // set in one-process-per-GPU mode with world > 1
  int cur = 0;                 // buffer holding the current state
  int host_y0 = 0;             // first lattice row of the caller's host planes (rank mode: the slab's)
  int vec = 4, tpb = 128, chunk = 128, pad = 0;
  bool use_graph = true;
  bool use_pdl = false;        // programmatic dependent launch between consecutive steps (1 GPU)
  int fuse_mode = -1;          // LBM_FUSE: 2 = two timesteps per pass, 1 = one, -1 = by slab size
  double last_ms = 0;
  long long last_launches = 0;
  std::string config;
};

namespace {

using lbm::StepArgs;

template <int VEC, int TPB>
cudaError_t launch_step_t(const StepArgs& a, int nblk, cudaStream_t st, bool pdl)
{
  if (!pdl) {
    lbm::lbm_step_kernel<VEC, TPB><<<nblk, TPB, 0, st>>>(a);
    return cudaGetLastError();
  }
  // programmatic dependent launch on the previous kernel of the stream (see the kernel prologue)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)nblk);
  cfg.blockDim = dim3(TPB);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, lbm::lbm_step_kernel<VEC, TPB>, a);
}

cudaError_t launch_step(int vec, int tpb, const StepArgs& a, int nblk, cudaStream_t st, bool pdl = false)
{
#define LBM_CASE(V, T) if (vec == V && tpb == T) return launch_step_t<V, T>(a, nblk, st, pdl);
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
  if (a.nxv > 1) {            // magic-number division by nxv, exact for dividends < 2^31
    int l = 0;
    while ((1u << l) < (unsigned)a.nxv) l++;
    const unsigned long long pw = 1ull << (31 + l);
    a.div_mul = (unsigned)((pw + (unsigned)a.nxv - 1) / (unsigned)a.nxv);
    a.div_shift = l - 1;
  }
  a.omega = h->p.omega;
  a.a1 = h->a1;
  a.a2 = h->a2;
  a.fuse_accel = fuse;
  for (int i = 0; i < 3; i++) {
    a.ghost_lo[i] = s.ghost_lo[cur ^ 1][i];
    a.ghost_hi[i] = s.ghost_hi[cur ^ 1][i];
  }
  a.partials = s.partials + (long long)slot * s.np;
  a.ps_dst = s.ps;
  a.dst_delta = 0;
  a.push = 3;
  if (h->comm && h->comm->halo == HALO_P2P && h->comm->in_kernel) {
    const Comm* c = h->comm;
    const long long last_row_first_item = (long long)(s.rows - 1) * a.nxv;
    a.ring_in = s.sync;
    a.ring_out_lo = c->peer_lo_flag;
    a.ring_out_hi = c->peer_hi_flag;
    a.ring_tickets = s.sync + 4;
    a.ring_timeout = s.sync + 2;
    a.ring_step = c->steps_done;
    a.nb_hi = s.nblk - (int)(last_row_first_item / h->tpb);
    a.nb_lo = (a.nxv + h->tpb - 1) / h->tpb;
    a.rot = a.nb_hi;
  }
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

int sync_all(lbm_lattice* h)
{
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    CK(cudaStreamSynchronize(s.stream));
  }
  return 0;
}

// all ranks have reached this point and their streams are idle (multi-process mode only)
int comm_barrier(lbm_lattice* h)
{
  if (!h->comm) return 0;
  Slab& s = h->slabs[0];
  NcclApi* n = nccl_api();
  CK(cudaStreamSynchronize(s.stream));
  NK(n->AllReduce(h->comm->scratch64, h->comm->scratch64, 1, ncclInt64, ncclSum, h->comm->nccl, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return 0;
}

// NCCL flavour of the halo: my first row's 4,7,8 go down, my last row's 2,5,6 go up, straight
// from / into the plane rows of buffer `b` (each a contiguous run of nx floats)
int nccl_halo_exchange(lbm_lattice* h, Slab& s, int b)
{
  NcclApi* n = nccl_api();
  Comm* c = h->comm;
  const int nx = h->p.nx;
  const int lo = (c->rank + c->world - 1) % c->world, hi = (c->rank + 1) % c->world;
  float* base = s.buf[b];
  NK(n->GroupStart());
  for (int i = 0; i < 3; i++) {
    NK(n->Send(base + LO_PLANES[i] * s.ps + (long long)nx, nx, ncclFloat, lo, c->nccl, s.stream));
    NK(n->Send(base + HI_PLANES[i] * s.ps + (long long)s.rows * nx, nx, ncclFloat, hi, c->nccl, s.stream));
    NK(n->Recv(base + LO_PLANES[i] * s.ps + (long long)(s.rows + 1) * nx, nx, ncclFloat, hi, c->nccl, s.stream));
    NK(n->Recv(base + HI_PLANES[i] * s.ps, nx, ncclFloat, lo, c->nccl, s.stream));
  }
  NK(n->GroupEnd());
  return 0;
}

// after the resident state changed from outside (upload / init): fill every ghost row
int refresh_ghosts(lbm_lattice* h)
{
  if (sync_all(h)) return 1;
  if (comm_barrier(h)) return 1;
  const int nx = h->p.nx;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    if (h->comm && h->comm->halo == HALO_NCCL) {
      if (nccl_halo_exchange(h, s, h->cur)) return 1;
      continue;
    }
    const int c = h->cur;
    lbm::halo_push_kernel<<<(nx + 255) / 256, 256, 0, s.stream>>>(
        s.buf[c], s.ps, nx, s.rows, s.ghost_lo[c][0], s.ghost_lo[c][1], s.ghost_lo[c][2],
        s.ghost_hi[c][0], s.ghost_hi[c][1], s.ghost_hi[c][2]);
    CK(cudaGetLastError());
  }
  if (sync_all(h)) return 1;
  return comm_barrier(h);
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
    CK(launch_step(h->vec, h->tpb, a, s.nblk, s.stream, h->use_pdl && i > 0));
    c ^= 1;
  }
  lbm::reduce_partials_kernel<<<h->chunk, 256, 0, s.stream>>>(s.partials, s.np, s.nblk, s.nblk, s.totals, s.counter);
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

// ---- two-step passes (LBM_FUSE=2) ---------------------------------------------------------------
// Two-step passes halve the DRAM traffic but cost ~20 % more instructions and run at lower
// occupancy: they win once a slab streams from HBM (+17 % at 16384^2, profiles/r1_tuning.md) and lose
// when it lives in L2 (1024^2: -40 %), so "auto" turns them on from 8 M cells (0.6 GB of state) up.
bool fused_ok(const lbm_lattice* h)
{
  // decided from global quantities only: every rank of a ring must take the same path
  const int min_rows = h->p.ny / h->world;
  if (h->fuse_mode == 1 || h->slabs.size() != 1 || h->vec != 4 || h->p.nx < 128 || min_rows < 4) return false;
  if (h->comm && (h->comm->halo != HALO_P2P || h->comm->per_step_allreduce)) return false;
  if (h->fuse_mode == 2) return true;
  return (long long)min_rows * h->p.nx >= (8LL << 20);
}

// t -> t+2 on rows 2..rows-1 of buffer cur^1, t+1 boundary rows into the strips (slots: 2 steps)
// tile shapes of the two-step kernel: {threads, rows relaxed to t+1, min blocks per SM}
struct FusedCfg { int tpb, ra, minb; };
// {224,14,3} measured best on B200 (profiles/r1_tuning.md section 7): 7 warps x 2 rows, 64.5 KB of shared
// memory, 80 registers, 3 blocks per SM (193 KB of shared memory in use leaves the L1 some room)
const FusedCfg FUSED_CFGS[] = {{224, 14, 3}, {256, 16, 2}, {288, 18, 2}, {320, 20, 2}, {192, 12, 3}, {160, 10, 4}};
int g_fused_cfg = 0;

template <int TPB, int RA, int MINB>
int launch_fused_t(const lbm::FusedArgs& a, int ntiles, cudaStream_t st)
{
  // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
  static bool configured[64] = {};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    CK(cudaFuncSetAttribute(lbm::lbm_fused2_kernel<TPB, RA, MINB>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, lbm::fused_smem(RA)));
    configured[dev] = true;
  }
  lbm::lbm_fused2_kernel<TPB, RA, MINB><<<ntiles, TPB, lbm::fused_smem(RA), st>>>(a);
  CK(cudaGetLastError());
  return 0;
}

int launch_fused(lbm_lattice* h, Slab& s, int cur, int fuse_b, int slot)
{
  lbm::FusedArgs a{};
  a.src = s.buf[cur];
  a.dst = s.buf[cur ^ 1];
  a.flags = s.flags;
  a.ps = s.ps;
  a.nx = h->p.nx;
  a.rows = s.rows;
  a.tiles_x = s.tiles_x;
  a.tiles_y = s.tiles_y;
  a.omega = h->p.omega;
  a.a1 = h->a1;
  a.a2 = h->a2;
  a.fuse_b = fuse_b;
  a.strip_lo = s.strip_lo;
  a.strip_hi = s.strip_hi;
  a.pse = s.pse;
  for (int i = 0; i < 3; i++) { a.nb_lo[i] = s.nb_lo[i]; a.nb_hi[i] = s.nb_hi[i]; }
  a.partials_a = s.partials + (long long)slot * s.np;
  a.partials_b = s.partials + (long long)(slot + 1) * s.np;
  if (h->comm && h->comm->in_kernel) {
    const Comm* c = h->comm;
    a.ring_in = s.sync;
    a.ring_out_lo = c->peer_lo_flag;
    a.ring_out_hi = c->peer_hi_flag;
    a.ring_tickets = s.sync + 4;
    a.ring_timeout = s.sync + 2;
    a.ring_step = c->steps_done;
    a.rot = s.tiles_x;
  }
  const int nt = s.tiles_x * s.tiles_y;
  switch (g_fused_cfg) {
    case 1: return launch_fused_t<256, 16, 2>(a, nt, s.stream);
    case 2: return launch_fused_t<288, 18, 2>(a, nt, s.stream);
    case 3: return launch_fused_t<320, 20, 2>(a, nt, s.stream);
    case 4: return launch_fused_t<192, 12, 3>(a, nt, s.stream);
    case 5: return launch_fused_t<160, 10, 4>(a, nt, s.stream);
    default: return launch_fused_t<224, 14, 3>(a, nt, s.stream);
  }
}

// rows 1 and `rows` of time t+2 from the strips (which now hold the neighbours' t+1 rows too):
// two launches of the ordinary step kernel on 3-row mini slabs; they also push the t+2 ghost rows
int launch_fixups(lbm_lattice* h, Slab& s, int cur, int fuse_b, int slot)
{
  const int nx = h->p.nx, nxv = nx / h->vec;
  const int nb_strip = (nxv + h->tpb - 1) / h->tpb;
  const int ntiles = s.tiles_x * s.tiles_y;
  for (int hi = 0; hi < 2; hi++) {
    StepArgs a = make_args(h, s, cur, fuse_b, slot);
    a.src = hi ? s.strip_hi : s.strip_lo;
    a.ps = s.pse;
    a.ps_dst = s.ps;
    a.rows = 1;
    a.nvec = nxv;
    a.dst_delta = hi ? (long long)(s.rows - 1) * nx : 0;
    a.push = hi ? 2 : 1;
    a.partials = s.partials + (long long)slot * s.np + ntiles + hi * nb_strip;
    a.rot = 0;                // ring ordering (if done in the kernel): this strip's side only, see `push`
    a.nb_lo = a.nb_hi = nb_strip;
    CK(launch_step(h->vec, h->tpb, a, nb_strip, s.stream));
  }
  return 0;
}

int run_impl(lbm_lattice* h, int iters, double* av_out)
{
  h->last_ms = 0;
  h->last_launches = 0;
  if (iters < 0) return fail("lbm_run: negative iteration count");
  if (iters == 0) return 0;
  const size_t nslab = h->slabs.size();
  Comm* comm = h->comm;
  const long long totals_before = h->slabs[0].totals_cap;
  if (ensure_totals(h, iters)) return 1;
  if (h->slabs[0].totals_cap != totals_before)
    for (auto& s : h->slabs) drop_graphs(s);   // the captured reduce node holds the old pointer
  // every rank is idle and out of any other API call before peers start writing ghost rows
  if (comm_barrier(h)) return 1;

  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    CK(cudaMemsetAsync(s.counter, 0, sizeof(long long), s.stream));
    CK(cudaEventRecord(s.ev_begin, s.stream));
    if (launch_accelerate(h, s, h->cur)) return 1;
  }

  int remaining = iters;
  int cur = h->cur;
  // ---- two timesteps per pass while at least two remain (LBM_FUSE=2)
  if (fused_ok(h)) {
    Slab& s = h->slabs[0];
    const int ntiles = s.tiles_x * s.tiles_y;
    const int nb_strip = ((h->p.nx / h->vec) + h->tpb - 1) / h->tpb;
    while (remaining >= 2) {
      const int pairs = std::min(remaining / 2, h->chunk / 2);
      for (int j = 0; j < pairs; j++) {
        const int fuse_b = (remaining - 2 * j - 2) > 0;
        const bool launches_order = comm && !comm->in_kernel;   // else the kernels order themselves
        if (launches_order) {
          lbm::wait_neighbours_kernel<<<1, 2, 0, s.stream>>>(s.sync, comm->steps_done, s.sync + 2);
          CK(cudaGetLastError());
        }
        if (launch_fused(h, s, cur, fuse_b, 2 * j)) return 1;
        if (launches_order) {   // my t+1 boundary rows are in the neighbours' strips; wait for theirs
          lbm::signal_neighbours_kernel<<<1, 2, 0, s.stream>>>(comm->peer_lo_flag, comm->peer_hi_flag,
                                                               comm->steps_done + 1);
          CK(cudaGetLastError());
          lbm::wait_neighbours_kernel<<<1, 2, 0, s.stream>>>(s.sync, comm->steps_done + 1, s.sync + 2);
          CK(cudaGetLastError());
        }
        if (comm) comm->steps_done++;
        if (launch_fixups(h, s, cur, fuse_b, 2 * j + 1)) return 1;
        if (launches_order) {
          lbm::signal_neighbours_kernel<<<1, 2, 0, s.stream>>>(comm->peer_lo_flag, comm->peer_hi_flag,
                                                               comm->steps_done + 1);
          CK(cudaGetLastError());
        }
        if (comm) comm->steps_done++;
        h->last_launches += launches_order ? 7 : 3;
        cur ^= 1;
      }
      lbm::reduce_partials_kernel<<<2 * pairs, 256, 0, s.stream>>>(s.partials, s.np, ntiles,
                                                                   ntiles + 2 * nb_strip, s.totals, s.counter);
      CK(cudaGetLastError());
      lbm::advance_counter_kernel<<<1, 1, 0, s.stream>>>(s.counter, 2 * pairs);
      CK(cudaGetLastError());
      h->last_launches += 2;
      remaining -= 2 * pairs;
    }
  }

  if (nslab == 1 && !comm && h->use_graph) {
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
  long long step_no = 0;     // steps launched one by one in this run (graph chunks never coexist with them)
  if (comm && comm->per_step_allreduce && comm->step_sums_cap < iters) {
    if (comm->step_sums) CK(cudaFree(comm->step_sums));
    comm->step_sums = nullptr;
    CK(cudaMalloc(&comm->step_sums, sizeof(double) * h->slabs[0].totals_cap));
    comm->step_sums_cap = h->slabs[0].totals_cap;
  }
  while (remaining > 0) {
    const int n = std::min(remaining, h->chunk);
    for (int i = 0; i < n; i++, step_no++) {
      const int fuse = (remaining - i) > 1;
      for (size_t k = 0; k < nslab; k++) {
        Slab& s = h->slabs[k];
        CK(cudaSetDevice(s.device));
        // my ring neighbours must have finished the previous step: their stores into my ghost
        // rows are complete and they no longer read the ghost rows I am about to overwrite
        if (nslab > 1 && step_no > 0) {
          const Slab& lo = h->slabs[(k + nslab - 1) % nslab];
          const Slab& hi = h->slabs[(k + 1) % nslab];
          CK(cudaStreamWaitEvent(s.stream, lo.ev_step[(step_no - 1) & 1], 0));
          CK(cudaStreamWaitEvent(s.stream, hi.ev_step[(step_no - 1) & 1], 0));
        }
        if (comm && comm->halo == HALO_P2P && !comm->in_kernel) {
          lbm::wait_neighbours_kernel<<<1, 2, 0, s.stream>>>(s.sync, comm->steps_done, s.sync + 2);
          CK(cudaGetLastError());
          h->last_launches++;
        }
        const StepArgs a = make_args(h, s, cur, fuse, i);
        CK(launch_step(h->vec, h->tpb, a, s.nblk, s.stream, h->use_pdl && nslab == 1 && !comm && i > 0));
        if (nslab > 1) CK(cudaEventRecord(s.ev_step[step_no & 1], s.stream));
        if (comm && comm->halo == HALO_P2P && !comm->in_kernel) {
          lbm::signal_neighbours_kernel<<<1, 2, 0, s.stream>>>(comm->peer_lo_flag, comm->peer_hi_flag,
                                                               comm->steps_done + 1);
          CK(cudaGetLastError());
          h->last_launches++;
        } else if (comm && comm->halo == HALO_NCCL) {
          if (nccl_halo_exchange(h, s, cur ^ 1)) return 1;
          h->last_launches++;
        }
      }
      if (comm && comm->per_step_allreduce) {
        // the north-star formulation: this step's slab total -> all ranks, right away
        Slab& s = h->slabs[0];
        const long long idx = step_no;
        lbm::reduce_partials_kernel<<<1, 256, 0, s.stream>>>(s.partials + (long long)i * s.np, s.np, s.nblk,
                                                            s.nblk, s.totals + idx, comm->scratch64);
        CK(cudaGetLastError());
        CK(cudaEventRecord(comm->ev_side, s.stream));
        CK(cudaStreamWaitEvent(comm->side, comm->ev_side, 0));
        NK(nccl_api()->AllReduce(s.totals + idx, comm->step_sums + idx, 1, ncclDouble, ncclSum,
                                 comm->nccl, comm->side));
        h->last_launches += 2;
      }
      if (comm) comm->steps_done++;
      h->last_launches += (long long)nslab;
      cur ^= 1;
    }
    if (!(comm && comm->per_step_allreduce)) {
      for (auto& s : h->slabs) {
        CK(cudaSetDevice(s.device));
        lbm::reduce_partials_kernel<<<n, 256, 0, s.stream>>>(s.partials, s.np, s.nblk, s.nblk, s.totals, s.counter);
        CK(cudaGetLastError());
        lbm::advance_counter_kernel<<<1, 1, 0, s.stream>>>(s.counter, n);
        CK(cudaGetLastError());
      }
      h->last_launches += 2 * (long long)nslab;
    }
    remaining -= n;
  }
  h->cur = cur;

  if (comm && comm->per_step_allreduce) {      // the run is over when its last allreduce is
    CK(cudaEventRecord(comm->ev_side, comm->side));
    CK(cudaStreamWaitEvent(h->slabs[0].stream, comm->ev_side, 0));
  }
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

  if (comm) {
    unsigned timed_out = 0;
    CK(cudaMemcpy(&timed_out, h->slabs[0].sync + 2, sizeof timed_out, cudaMemcpyDeviceToHost));
    if (timed_out) return fail("lbm_run: timed out waiting for a neighbour GPU (rank %d)", comm->rank);
  }

  std::fill(av_out, av_out + iters, 0.0);
  std::vector<double> tmp((size_t)iters);
  if (!comm) {
    for (auto& s : h->slabs) {          // fixed slab order: the cross-GPU sum is deterministic
      CK(cudaSetDevice(s.device));
      CK(cudaMemcpy(tmp.data(), s.totals, sizeof(double) * iters, cudaMemcpyDeviceToHost));
      for (int t = 0; t < iters; t++) av_out[t] += tmp[t];
    }
  } else if (comm->per_step_allreduce) {
    Slab& s = h->slabs[0];
    CK(cudaStreamSynchronize(comm->side));
    CK(cudaMemcpy(av_out, comm->step_sums, sizeof(double) * iters, cudaMemcpyDeviceToHost));
    (void)s;
  } else {
    // all-gather every rank's per-step totals, then add them in rank order on the host
    Slab& s = h->slabs[0];
    double* gathered = nullptr;
    CK(cudaMalloc(&gathered, sizeof(double) * (size_t)iters * comm->world));
    const ncclResult_t nr = nccl_api()->AllGather(s.totals, gathered, (size_t)iters, ncclDouble,
                                                  comm->nccl, s.stream);
    cudaError_t ce = nr == ncclSuccess ? cudaStreamSynchronize(s.stream) : cudaErrorUnknown;
    for (int r = 0; r < comm->world && ce == cudaSuccess; r++) {
      ce = cudaMemcpy(tmp.data(), gathered + (size_t)r * iters, sizeof(double) * iters, cudaMemcpyDeviceToHost);
      for (int t = 0; t < iters; t++) av_out[t] += tmp[t];
    }
    cudaFree(gathered);
    if (nr != ncclSuccess) return fail("NCCL error gathering av_vels: %s", nccl_api()->GetErrorString(nr));
    if (ce != cudaSuccess) return fail("CUDA error gathering av_vels: %s", cudaGetErrorString(ce));
  }
  const double denom = (double)h->tot_cells;
  for (int t = 0; t < iters; t++) av_out[t] /= denom;
  return 0;
}


void read_tuning(lbm_lattice* h)
{
  const int nx = h->p.nx;
  int vec = (nx % 4 == 0) ? 4 : (nx % 2 == 0) ? 2 : 1;
  const int want_vec = env_int("LBM_VEC", vec);
  if ((want_vec == 1 || want_vec == 2 || want_vec == 4) && nx % want_vec == 0) vec = want_vec;
  h->vec = vec;
  // 128-thread blocks: best or within 1 % of best on B200 for both the HBM-streaming and the
  // L2-resident regime (profiles/r1_tuning.md)
  const int want_tpb = env_int("LBM_TPB", 128);
  h->tpb = (want_tpb == 64 || want_tpb == 256 || want_tpb == 512) ? want_tpb : 128;
  h->chunk = std::max(2, env_int("LBM_CHUNK", 128)) & ~1;
  h->use_graph = env_int("LBM_GRAPH", 1) != 0;
  h->use_pdl = env_int("LBM_PDL", -1) != 0;   // -1 = decide per slab size (create_slab)
  h->pad = std::max(0, env_int("LBM_PLANE_PAD", 0));
  h->fuse_mode = env_int("LBM_FUSE", -1);
  g_fused_cfg = std::min(5, std::max(0, env_int("LBM_FUSE_CFG", 0)));
}

// device objects of one slab; obstacles_rows points at the slab's first row
int create_slab(lbm_lattice* h, Slab& s, const int* obstacles_rows, long long* fluid_cells)
{
  const int nx = h->p.nx, ny = h->p.ny;
  const long long cells = (long long)(s.rows + 2) * nx;
  s.ps = plane_stride(s.rows, nx, h->pad);
  s.nvec = (long long)s.rows * (nx / h->vec);
  if (s.nvec >= (1LL << 31)) return fail("lbm_create: slab too large for 32-bit work index");
  s.nblk = (int)((s.nvec + h->tpb - 1) / h->tpb);
  // PDL pays once a step is longer than a launch: measured +2 % at 1024^2 (2048 blocks), +10 % at
  // 128^2 but -50 % at 256^2 inside graphs (profiles/r1_tuning.md), so auto = multi-wave grids only
  if (env_int("LBM_PDL", -1) < 0) h->use_pdl = s.nblk >= 148 * 8;
  s.owns_accel_row = (ny - 2 >= s.y0 && ny - 2 < s.y0 + s.rows);
  s.accel_row = ny - 2 - s.y0 + 1;

  CK(cudaSetDevice(s.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, s.device));
  if (prop.major < 10)
    return fail("device %d (%s) is not an sm_100-class GPU; this library is built for sm_100a only",
                s.device, prop.name);
  CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&s.ev_begin));
  CK(cudaEventCreate(&s.ev_end));
  CK(cudaEventCreateWithFlags(&s.ev_step[0], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s.ev_step[1], cudaEventDisableTiming));
  CK(cudaMalloc(&s.base, slab_bytes(s.ps, nx)));
  CK(cudaMemset(s.base, 0, slab_bytes(s.ps, nx)));
  s.buf[0] = reinterpret_cast<float*>(s.base);
  s.buf[1] = s.buf[0] + 9 * s.ps;
  s.sync = reinterpret_cast<unsigned*>(s.buf[1] + 9 * s.ps);
  s.pse = strip_stride(nx);
  s.strip_lo = strip_of(s.base, s.ps, nx, 0);
  s.strip_hi = strip_of(s.base, s.ps, nx, 1);
  s.tiles_x = (nx + lbm::F_TX - 1) / lbm::F_TX;
  {
    const int ty = FUSED_CFGS[g_fused_cfg].ra - 2;
    s.tiles_y = (std::max(s.rows - 2, 0) + ty - 1) / ty;
  }
  {
    const int nb_strip = ((nx / h->vec) + h->tpb - 1) / h->tpb;
    s.np = std::max(s.nblk, s.tiles_x * s.tiles_y + 2 * nb_strip);
  }
  CK(cudaMalloc(&s.flags, (size_t)cells));
  CK(cudaMalloc(&s.partials, sizeof(double) * (size_t)h->chunk * s.np));
  CK(cudaMalloc(&s.counter, sizeof(long long)));
  CK(cudaMemset(s.counter, 0, sizeof(long long)));

  // flags: bit 0 obstacle, bit 1 fluid cell of the accelerated row (global ny-2)
  std::vector<uint8_t> fl((size_t)cells, 0);
  long long fluid = 0;
  for (int r = 0; r < s.rows; r++) {
    const int gy = s.y0 + r;
    const int* orow = obstacles_rows + (size_t)r * nx;
    uint8_t* frow = fl.data() + (size_t)(r + 1) * nx;
    for (int x = 0; x < nx; x++) {
      const bool ob = orow[x] != 0;
      frow[x] = (uint8_t)((ob ? lbm::FLAG_OBSTACLE : 0) | ((!ob && gy == ny - 2) ? lbm::FLAG_ACCEL : 0));
      fluid += !ob;
    }
  }
  *fluid_cells += fluid;
  CK(cudaMemcpy(s.flags, fl.data(), (size_t)cells, cudaMemcpyHostToDevice));
  return 0;
}

// ghost destinations of slab `s` inside the allocations `lo_base` / `hi_base` of its ring
// neighbours, whose geometry is (lo_rows, lo_ps) / hi_ps
void wire_ghosts(Slab& s, int nx, char* lo_base, int lo_rows, long long lo_ps, char* hi_base,
                 long long hi_ps)
{
  for (int b = 0; b < 2; b++) {
    float* lo_buf = reinterpret_cast<float*>(lo_base) + (long long)b * 9 * lo_ps;
    float* hi_buf = reinterpret_cast<float*>(hi_base) + (long long)b * 9 * hi_ps;
    for (int i = 0; i < 3; i++) {
      s.ghost_lo[b][i] = lo_buf + LO_PLANES[i] * lo_ps + (long long)(lo_rows + 1) * nx;
      s.ghost_hi[b][i] = hi_buf + HI_PLANES[i] * hi_ps;
    }
  }
  const long long pse = strip_stride(nx);
  for (int i = 0; i < 3; i++) {
    s.nb_lo[i] = strip_of(lo_base, lo_ps, nx, 1) + LO_PLANES[i] * pse + 2LL * nx;
    s.nb_hi[i] = strip_of(hi_base, hi_ps, nx, 0) + HI_PLANES[i] * pse;
  }
}

void set_config_string(lbm_lattice* h)
{
  char cfg[320];
  const char* red = (h->comm && h->comm->per_step_allreduce) ? " reduce=allreduce-per-step" : "";
  const char* mode = h->comm ? (h->comm->halo == HALO_NCCL ? "ranks+nccl-sendrecv"
                                : h->comm->in_kernel ? "ranks+ipc-peer-stores+in-kernel-ring"
                                                     : "ranks+ipc-peer-stores+wait/signal-kernels")
                             : (h->slabs.size() > 1 ? "one-process+peer-stores" : "single-gpu");
  snprintf(cfg, sizeof cfg, "vec=%d tpb=%d chunk=%d graph=%d pdl=%d fuse=%d slabs=%d halo=%s%s plane_stride=%lld",
           h->vec, h->tpb, h->chunk, (int)(h->use_graph && h->world == 1),
           (int)(h->use_pdl && h->world == 1), fused_ok(h) ? 2 : 1, h->world, mode, red,
           h->slabs[0].ps);
  h->config = cfg;
}

int common_checks(lbm_lattice** out, const lbm_params* p, const int* obstacles)
{
  if (!out || !p || !obstacles) return fail("lbm_create: null argument");
  *out = nullptr;
  if (p->nx < 1 || p->ny < 1) return fail("lbm_create: bad grid %d x %d", p->nx, p->ny);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    return fail("lbm_create: no CUDA device available (this library has no CPU fallback)");
  }
  return 0;
}

lbm_lattice* new_lattice(const lbm_params* p, int world)
{
  lbm_lattice* h = new lbm_lattice();
  h->p = *p;
  // kernels.cl:17-18: float product, double divide, rounded to float
  h->a1 = (float)((double)(p->density * p->accel) / 9.0);
  h->a2 = (float)((double)(p->density * p->accel) / 36.0);
  h->world = world;
  read_tuning(h);
  return h;
}

int create_impl(lbm_lattice** out, const lbm_params* p, const int* obstacles, int first_device,
                int nslab)
{
  if (common_checks(out, p, obstacles)) return 1;
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (nslab < 1 || first_device < 0 || first_device + nslab > ndev)
    return fail("lbm_create: %d GPU(s) requested from device %d but %d visible", nslab, first_device, ndev);
  if (nslab > 1 && p->ny / nslab < 3)
    return fail("lbm_create: %d rows cannot be split into %d slabs of >= 3 rows", p->ny, nslab);

  lbm_lattice* h = new_lattice(p, nslab);
  const int nx = p->nx;
  h->slabs.resize(nslab);
  for (int k = 0; k < nslab; k++) {
    Slab& s = h->slabs[k];
    s.device = first_device + k;
    s.rank = k;
    lbm_slab_rows(p->ny, nslab, k, &s.y0, &s.rows);
    if (create_slab(h, s, obstacles + (size_t)s.y0 * nx, &h->tot_cells)) { lbm_destroy(h); return 1; }
  }
  for (int k = 0; k < nslab; k++) {
    Slab& s = h->slabs[k];
    Slab& lo = h->slabs[(k + nslab - 1) % nslab];
    Slab& hi = h->slabs[(k + 1) % nslab];
    wire_ghosts(s, nx, lo.base, lo.rows, lo.ps, hi.base, hi.ps);
    if (nslab > 1) {
      cudaSetDevice(s.device);
      for (const Slab* nb : {&lo, &hi}) {
        if (nb->device == s.device) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, s.device, nb->device);
        if (!can) {
          fail("GPU %d cannot access GPU %d peer memory", s.device, nb->device);
          lbm_destroy(h);
          return 1;
        }
        cudaError_t e = cudaDeviceEnablePeerAccess(nb->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) {
          fail("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
          lbm_destroy(h);
          return 1;
        }
      }
    }
  }
  set_config_string(h);
  *out = h;
  return 0;
}

// one process per GPU: NCCL communicator, global cell count, IPC mapping of the two neighbours
int attach_comm(lbm_lattice* h, int rank, int world, const void* unique_id)
{
  NcclApi* n = nccl_api();
  if (!n) return fail("lbm_create_rank: libnccl.so.2 could not be loaded (%s)", dlerror());
  if (!unique_id) return fail("lbm_create_rank: world > 1 needs an ncclUniqueId");
  Slab& s = h->slabs[0];
  Comm* c = new Comm();
  h->comm = c;
  c->rank = rank;
  c->world = world;
  const char* halo = getenv("LBM_HALO");
  c->halo = (halo && !strcmp(halo, "nccl")) ? HALO_NCCL : HALO_P2P;
  const char* red = getenv("LBM_REDUCE");
  c->per_step_allreduce = red && !strcmp(red, "step");
  CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof id);
  NK(n->CommInitRank(&c->nccl, world, id, rank));
  CK(cudaMalloc(&c->scratch64, sizeof(long long) * 2));
  CK(cudaMemset(c->scratch64, 0, sizeof(long long) * 2));

  // global number of fluid cells (d2q9-bgk.c:146-152 counts them over the whole grid)
  long long* d_cnt = c->scratch64 + 1;
  CK(cudaMemcpy(d_cnt, &h->tot_cells, sizeof(long long), cudaMemcpyHostToDevice));
  NK(n->AllReduce(d_cnt, d_cnt, 1, ncclInt64, ncclSum, c->nccl, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaMemcpy(&h->tot_cells, d_cnt, sizeof(long long), cudaMemcpyDeviceToHost));

  const int nx = h->p.nx;
  const int lo = (rank + world - 1) % world, hi = (rank + 1) % world;
  int lo_y0, lo_rows, hi_y0, hi_rows;
  lbm_slab_rows(h->p.ny, world, lo, &lo_y0, &lo_rows);
  lbm_slab_rows(h->p.ny, world, hi, &hi_y0, &hi_rows);
  const long long lo_ps = plane_stride(lo_rows, nx, h->pad), hi_ps = plane_stride(hi_rows, nx, h->pad);

  if (c->halo == HALO_NCCL) {
    // ghost stores of the kernel land in a scratch row; the real halo moves by send/recv
    CK(cudaMalloc(&c->dummy_ghost, sizeof(float) * (size_t)nx));
    for (int b = 0; b < 2; b++)
      for (int i = 0; i < 3; i++) s.ghost_lo[b][i] = s.ghost_hi[b][i] = c->dummy_ghost;
    c->ready = true;
    return 0;
  }

  // exchange CUDA-IPC handles of the slab allocations through the communicator
  cudaIpcMemHandle_t mine;
  CK(cudaIpcGetMemHandle(&mine, s.base));
  char* d_all = nullptr;
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  CK(cudaMalloc(&d_all, hb * (world + 1)));
  CK(cudaMemcpy(d_all + hb * world, &mine, hb, cudaMemcpyHostToDevice));
  NK(n->AllGather(d_all + hb * world, d_all, hb, ncclChar, c->nccl, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  std::vector<cudaIpcMemHandle_t> all(world);
  CK(cudaMemcpy(all.data(), d_all, hb * world, cudaMemcpyDeviceToHost));
  CK(cudaFree(d_all));
  CK(cudaIpcOpenMemHandle(&c->peer_lo, all[lo], cudaIpcMemLazyEnablePeerAccess));
  if (hi == lo) c->peer_hi = c->peer_lo;
  else CK(cudaIpcOpenMemHandle(&c->peer_hi, all[hi], cudaIpcMemLazyEnablePeerAccess));

  wire_ghosts(s, nx, (char*)c->peer_lo, lo_rows, lo_ps, (char*)c->peer_hi, hi_ps);
  // my lower neighbour counts me as its UPPER neighbour (its sync[1]); the upper one as its LOWER
  c->peer_lo_flag = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(c->peer_lo) + 18 * lo_ps) + 1;
  c->peer_hi_flag = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(c->peer_hi) + 18 * hi_ps) + 0;
  // Ring ordering inside the step kernel needs every row to start on its own 128-byte line (an
  // interior block must not pull a stale copy of a ghost row's tail into L1 before the boundary
  // block has seen the neighbour's flag); otherwise two tiny wait/signal launches bracket each step.
  const char* ring = getenv("LBM_RING");
  c->in_kernel = (nx % 32 == 0) && !(ring && !strcmp(ring, "kernels"));
  if (comm_barrier(h)) return 1;
  c->ready = true;
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
  if (world == 1 && rank == 0) return create_impl(out, params, obstacles_slab, device, 1);
  if (common_checks(out, params, obstacles_slab)) return 1;
  if (world < 1 || rank < 0 || rank >= world) return fail("lbm_create_rank: rank %d of %d", rank, world);
  if (params->ny / world < 3)
    return fail("lbm_create_rank: %d rows cannot be split into %d slabs of >= 3 rows", params->ny, world);
  lbm_lattice* h = new_lattice(params, world);
  h->slabs.resize(1);
  Slab& s = h->slabs[0];
  s.device = device;
  s.rank = rank;
  lbm_slab_rows(params->ny, world, rank, &s.y0, &s.rows);
  h->host_y0 = s.y0;
  if (create_slab(h, s, obstacles_slab, &h->tot_cells) || attach_comm(h, rank, world, nccl_unique_id)) {
    lbm_destroy(h);
    return 1;
  }
  set_config_string(h);
  *out = h;
  return 0;
}

int lbm_comm_unique_id(void* out128)
{
  NcclApi* n = nccl_api();
  if (!n) return fail("lbm_comm_unique_id: libnccl.so.2 could not be loaded");
  if (!out128) return fail("null argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in the C ABI");
  ncclUniqueId id;
  NK(n->GetUniqueId(&id));
  memcpy(out128, &id, sizeof id);
  return 0;
}

void lbm_destroy(lbm_lattice* h)
{
  if (!h) return;
  for (auto& s : h->slabs) {
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
  }
  if (h->comm) {
    Comm* c = h->comm;
    NcclApi* n = nccl_api();
    if (c->nccl && n && c->ready) {        // nobody unmaps memory a peer may still write to
      Slab& s = h->slabs[0];
      n->AllReduce(c->scratch64, c->scratch64, 1, ncclInt64, ncclSum, c->nccl, s.stream);
      cudaStreamSynchronize(s.stream);
    }
    if (c->peer_hi && c->peer_hi != c->peer_lo) cudaIpcCloseMemHandle(c->peer_hi);
    if (c->peer_lo) cudaIpcCloseMemHandle(c->peer_lo);
    if (c->dummy_ghost) cudaFree(c->dummy_ghost);
    if (c->step_sums) cudaFree(c->step_sums);
    if (c->ev_side) cudaEventDestroy(c->ev_side);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->scratch64) cudaFree(c->scratch64);
    if (c->nccl && n) n->CommDestroy(c->nccl);
    delete c;
  }
  for (auto& s : h->slabs) {
    cudaSetDevice(s.device);
    drop_graphs(s);
    if (s.base) cudaFree(s.base);
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
  if (sync_all(h)) return 1;
  return comm_barrier(h);
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
  if (h->comm) {
    Slab& s = h->slabs[0];
    double* d = reinterpret_cast<double*>(h->comm->scratch64 + 1);
    CK(cudaMemcpy(d, &total, sizeof(double), cudaMemcpyHostToDevice));
    NK(nccl_api()->AllReduce(d, d, 1, ncclDouble, ncclSum, h->comm->nccl, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaMemcpy(&total, d, sizeof(double), cudaMemcpyDeviceToHost));
  }
  *av_vel = (float)(total / (double)h->tot_cells);
  return 0;
}

int lbm_total_density(lbm_lattice* h, double* total_out)
{
  if (!h || !total_out) return fail("null argument");
  double total = 0;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    const int nblk = std::min(s.nblk, 148 * 8);
    lbm::total_density_kernel<256><<<nblk, 256, 0, s.stream>>>(s.buf[h->cur], s.ps, h->p.nx, s.rows, s.partials);
    CK(cudaGetLastError());
    std::vector<double> part((size_t)nblk);
    CK(cudaMemcpyAsync(part.data(), s.partials, sizeof(double) * nblk, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (double v : part) total += v;
  }
  if (h->comm) {
    Slab& s = h->slabs[0];
    double* d = reinterpret_cast<double*>(h->comm->scratch64 + 1);
    CK(cudaMemcpy(d, &total, sizeof(double), cudaMemcpyHostToDevice));
    NK(nccl_api()->AllReduce(d, d, 1, ncclDouble, ncclSum, h->comm->nccl, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaMemcpy(&total, d, sizeof(double), cudaMemcpyDeviceToHost));
  }
  *total_out = total;
  return 0;
}

int lbm_macroscopic(lbm_lattice* h, float* ux, float* uy, float* speed, float* pressure)
{
  if (!h || !ux || !uy || !speed || !pressure) return fail("null argument");
  const int nx = h->p.nx;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    const size_t n = (size_t)s.rows * nx;
    float* scratch = nullptr;
    CK(cudaMalloc(&scratch, sizeof(float) * 4 * n));
    lbm::macroscopic_kernel<<<148 * 8, 256, 0, s.stream>>>(s.buf[h->cur], s.flags, s.ps, nx, s.rows,
                                                         h->p.density, scratch, scratch + n,
                                                         scratch + 2 * n, scratch + 3 * n);
    CK(cudaGetLastError());
    float* outs[4] = {ux, uy, speed, pressure};
    for (int k = 0; k < 4; k++)
      CK(cudaMemcpyAsync(outs[k] + (size_t)(s.y0 - h->host_y0) * nx, scratch + k * n, sizeof(float) * n,
                         cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaFree(scratch));
  }
  return 0;
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
