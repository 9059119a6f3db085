// lbm_engine.cu -- host side of the C ABI in include/lbm_b200.h: owns the device memory, streams,
// CUDA graphs, TMA descriptors, the slab decomposition and the inter-GPU plumbing, and launches the
// kernels of lbm_kernels.cuh / lbm_stream.cuh.
//
// What it replaces in the reference (d2q9-bgk.c): the t_ocl object bundle (:35-67), its creation
// (:642-780), upload (:159-201), the `for tt` loop with its buffer ping-pong (:203-234),
// timestep/accelerate_flow/comp_func host wrappers (:294-424) including the per-step clFinish +
// 4*nx*ny-byte read-back + serial host sum (:408-423), download (:237-272) and release (:803-809).
//
// Design notes
//  * A lattice is a ring of row slabs, one per GPU.  Slab storage has GHOST (4) ghost rows below and
//    above the owned rows, all nine planes, plus the obstacle flags of those rows: copies of the
//    neighbouring slab's boundary rows (the slab's own opposite edge when the ring has one member,
//    which is how the y-periodic wrap works).  Whoever computes a boundary row also stores it into
//    the neighbour's ghost zone -- a peer pointer over NVLink when there are several GPUs (direct
//    peer access when one process drives all GPUs, CUDA-IPC mappings when there is one process per
//    GPU).  Slabs are ordered only against their two ring neighbours, by release/acquire counters
//    that the boundary blocks of the kernels themselves wait on and publish; never against the
//    host, and with no collective on the data path.  LBM_HALO=nccl (one process per GPU, one-step
//    kernel) is the library-only contrast.
//  * Time loop: slabs that stream from HBM advance S timesteps per launch with the TMA/mbarrier
//    streaming kernel (lbm_stream.cuh); small or odd-shaped slabs, and the last iters mod S steps,
//    use the one-step kernel (lbm_kernels.cuh).
//  * Between API calls the resident state is always the reference's canonical post-step state.
//    Inside lbm_run the inflow acceleration of step t+1 is folded into the store epilogue of step t;
//    the first step of a run is preceded by a stand-alone accelerate kernel and the last step of a
//    run does not pre-accelerate.  Every run starts by refreshing the ghost zones.
//  * The average-velocity reduction never leaves the device during a run: per-block double sums per
//    step, reduced per chunk of steps by a second kernel into a per-step totals array.  After the
//    last step the per-slab totals are combined in slab order, so the result does not depend on
//    timing.  LBM_REDUCE=step: the last block of every launch reduces the step's partials and
//    pushes the slab total to every rank over NVLink -- the north-star's per-step allreduce, done by
//    the step kernel itself.
//  * On one GPU whole chunks of one-step launches are replayed from a CUDA graph to take the launch
//    overhead out of small grids.
#include <cuda.h>
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
#include "lbm_stream.cuh"
#include "lbm_resident.cuh"

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

bool env_is(const char* name, const char* value)
{
  const char* v = getenv(name);
  return v && !strcmp(v, value);
}

// ---- NCCL, bound at run time ------------------------------------------------------------------
// dlopen by soname: inside a PyTorch process this resolves to the libnccl.so.2 torch already
// loaded, in the plain C program to the system one.  Only the one-process-per-GPU mode needs it,
// and only for set-up (unique id -> communicator, IPC handle exchange) and barriers between API
// calls; the time loop itself uses NCCL only with LBM_HALO=nccl.
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

// ---- TMA descriptors: cuTensorMapEncodeTiled through the runtime's driver entry point (the
// library does not link libcuda) ----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (tried) return fn;
  tried = true;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    fn = reinterpret_cast<EncodeTiledFn>(p);
  else
    cudaGetLastError();
  return fn;
}

enum { HALO_P2P = 0, HALO_NCCL = 1 };
constexpr int G = lbm::GHOST;
constexpr int MAX_WORLD = 64;

// one-process-per-GPU plumbing of a slab
struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
  int halo = HALO_P2P;
  std::vector<void*> peers;      // IPC mappings of every other rank's slab allocation (peers[rank] = null)
  float* dummy_ghost = nullptr;  // NCCL halo: the kernel's ghost stores go nowhere useful
  long long* scratch64 = nullptr;
  bool ready = false;            // fully attached: destroy may run its closing barrier
};

// tile shapes of the streaming kernel: {timesteps per pass, warps (= rows per batch) per group,
// TMA stages, blocks per SM}
struct StreamCfg { int s, nw, k0, minb; };
// measured on B200, 16384^2 (profiles/r2_tuning.md): 0: 165.5, 5: 162.1, 2: 159.4, 3: 152.0, 1: 150.1, 6: 140.3 GLUPS
const StreamCfg STREAM_CFGS[] = {
    {2, 4, 3, 2},   // 0: 288 threads, 106 KB of shared memory: two blocks per SM (default)
    {3, 6, 3, 1},   // 1: 608 threads, 219 KB (default for LBM_FUSE=3)
    {2, 8, 3, 1},   // 2: 544 threads, 201 KB
    {4, 4, 2, 1},   // 3: 544 threads, 179 KB (default for LBM_FUSE=4)
    {2, 3, 2, 3},   // 4: 224 threads, 68 KB: three blocks per SM
    {2, 4, 2, 2},   // 5: as 0 with two TMA stages
    {3, 4, 2, 1},   // 6: 416 threads, 133 KB
};
constexpr int N_STREAM_CFGS = sizeof(STREAM_CFGS) / sizeof(STREAM_CFGS[0]);

struct Slab {
  int        device = 0;
  int        rank = 0;        // position in the ring
  int        y0 = 0, rows = 0;
  long long  ps = 0;          // plane stride (floats)
  char*      base = nullptr;  // one allocation: buffer 0 | buffer 1 | 256 B of sync words | flags
  float*     buf[2] = {nullptr, nullptr};
  // sync words: [0] phases finished by my lower neighbour's top edge, [1] by my upper neighbour's
  // bottom edge, [2] timeout flag, [4] [5] [6] tickets (lo edge, hi edge, whole-grid launches)
  unsigned*  sync = nullptr;
  uint8_t*   flags = nullptr;
  // ring neighbours as seen from this slab's device
  char*      lo_base = nullptr;
  char*      hi_base = nullptr;
  int        lo_rows = 0;
  long long  lo_ps = 0, hi_ps = 0;
  unsigned*  ring_out_lo = nullptr;   // lower neighbour's sync[1]
  unsigned*  ring_out_hi = nullptr;   // upper neighbour's sync[0]
  float*     ghost_lo[2][3] = {};  // one-step kernel: [buffer][plane 4,7,8] nearest ghost row above the lower neighbour
  float*     ghost_hi[2][3] = {};  //                  [buffer][plane 2,5,6] nearest ghost row below the upper neighbour
  float*     gz_lo[2] = {};        // plane 0 of the lower neighbour's upper ghost zone, per buffer
  float*     gz_hi[2] = {};        // plane 0 of the upper neighbour's lower ghost zone
  // LBM_REDUCE=step: every rank's per-step slab totals, [world][cap] doubles in each slab; peers' copies
  double*    allred = nullptr;
  double*    allred_peer[MAX_WORLD] = {};
  long long  allred_cap = 0;
  int        np = 0;               // partial-sum slots per step
  int        tiles_x = 0, tiles_y = 0, tile_h = 0, tall_rows = 0, tile_h2 = 0;
  CUtensorMap tm_state, tm_flags;       // tile boxes: 128 columns (flags: 144) x NW rows
  CUtensorMap tm_state_w, tm_flags_w;   // periodic-wrap boxes: 4 columns (flags: 16) x NW rows
  double*    partials = nullptr;   // [chunk][np]
  double*    totals = nullptr;     // [totals_cap] per-step speed totals of this slab
  long long* counter = nullptr;
  long long  totals_cap = 0;
  int        nblk = 0;
  long long  nvec = 0;
  float*     macro = nullptr;      // lbm_macroscopic scratch (4 planes), allocated on first use
  unsigned long long* inbox = nullptr;   // persistent small-lattice kernel: per-block halo inboxes ({tag : float} words)
  unsigned long long* trace = nullptr;   // LBM_STREAM_TRACE: per-tile timestamps of the latest streaming pass
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev_begin = nullptr, ev_end = nullptr;
  cudaGraphExec_t graph[2] = {nullptr, nullptr};
  bool       owns_accel_row = false;
  int        accel_row = 0;        // storage row of global row ny-2
};

// first / one-past-last output row (storage index) of row `by` of tiles of the streaming kernel
int tile_row0(const Slab& s, int by)
{
  return by < s.tall_rows ? G + s.tile_h * by : G + s.tile_h * s.tall_rows + s.tile_h2 * (by - s.tall_rows);
}
int tile_row1(const Slab& s, int by)
{
  return std::min(tile_row0(s, by) + (by < s.tall_rows ? s.tile_h : s.tile_h2), G + s.rows);
}

const int LO_PLANES[3] = {4, 7, 8};   // pulled by the row below (kernels.cl:94,97,98)
const int HI_PLANES[3] = {2, 5, 6};   // pulled by the row above (kernels.cl:92,95,96)

long long plane_stride(int rows, int nx, int pad)
{
  const long long cells = (long long)(rows + 2 * G) * nx;
  return ((cells + 31) / 32) * 32 + ((long long)pad / 32) * 32;
}

size_t flags_bytes(int rows, int nx) { return (((size_t)(rows + 2 * G) * nx + 255) / 256) * 256; }

// buffer 0 | buffer 1 | 256 B of sync words | flags
size_t slab_bytes(long long ps, int rows, int nx) { return sizeof(float) * 18 * (size_t)ps + 256 + flags_bytes(rows, nx); }
float* buf_of(char* base, long long ps, int b) { return reinterpret_cast<float*>(base) + (long long)b * 9 * ps; }
unsigned* sync_of(char* base, long long ps) { return reinterpret_cast<unsigned*>(reinterpret_cast<float*>(base) + 18 * ps); }
uint8_t* flags_of(char* base, long long ps) { return reinterpret_cast<uint8_t*>(base) + sizeof(float) * 18 * (size_t)ps + 256; }

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
  bool stream_pdl = true;      // ... between consecutive streaming passes (LBM_STREAM_PDL)
  int fuse_mode = -1;          // LBM_FUSE: S >= 2 timesteps per pass, 1 = one-step kernel only, -1 = by slab size
  int stream_cfg = 0;          // index into STREAM_CFGS (fixed at create: tiles and TMA boxes depend on it)
  int tile_h_max = 0;          // LBM_TILE_H: upper bound of the tile height (0 = default)
  bool ring = false;           // several slabs ordered by release/acquire counters in peer memory
  bool ring_in_kernel = false; // one-step kernel: counters handled by its boundary blocks (else wait/signal launches)
  bool reduce_per_step = false;// LBM_REDUCE=step
  bool stall_test = false;     // LBM_TEST_RING_STALL: this rank never publishes (negative test of the time-out)
  // small lattices on one GPU: the whole time loop as one persistent cooperative launch (lbm_resident.cuh)
  bool resident = false;
  int res_rows = 1, res_tpb = 512, res_nblk = 0, res_chunk = 1200;   // rows per block, threads, blocks, steps per launch
  int res_cpt = 1;             // cells per thread (kernel instantiation)
  unsigned res_base = 0;       // halo tags handed out so far (one per timestep)
  unsigned phase = 0;          // ring phases completed since creation (all slabs advance together)
  bool poisoned = false;       // a run failed half-way: the ring counters are out of step
  double last_ms = 0;
  long long last_launches = 0;
  std::string config;
};

namespace {

using lbm::StepArgs;
using lbm::StreamArgs;

constexpr int ALLRED_CAP = 4096;     // per-step totals kept per rank between two drains (LBM_REDUCE=step)

double* allred_of(char* base, long long ps, int rows, int nx)
{
  return reinterpret_cast<double*>(reinterpret_cast<char*>(flags_of(base, ps)) + flags_bytes(rows, nx));
}

size_t slab_bytes_total(long long ps, int rows, int nx, int world)
{
  return slab_bytes(ps, rows, nx) + sizeof(double) * (size_t)ALLRED_CAP * (size_t)world;
}

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

void fill_reduce(const lbm_lattice* h, const Slab& s, lbm::StepReduce& r, long long step_index)
{
  r = lbm::StepReduce{};
  if (!h->reduce_per_step) return;
  r.world = h->world;
  r.rank = s.rank;
  r.cap = ALLRED_CAP;
  r.slot = (int)(step_index % ALLRED_CAP);
  r.ticket = s.sync + 7;
  for (int k = 0; k < h->world; k++) r.peer[k] = s.allred_peer[k];
}

// the one-step kernel works on a view of the slab whose row 0 is the nearest lower ghost row
StepArgs make_args(const lbm_lattice* h, const Slab& s, int cur, int fuse, int slot, long long step_index)
{
  StepArgs a{};
  const long long view = (long long)(G - 1) * h->p.nx;
  a.src = s.buf[cur] + view;
  a.dst = s.buf[cur ^ 1] + view;
  a.flags = s.flags + view;
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
  if (h->ring && h->ring_in_kernel) {
    const long long last_row_first_item = (long long)(s.rows - 1) * a.nxv;
    a.ring_in = s.sync;
    a.ring_out_lo = h->stall_test ? s.sync + 8 : s.ring_out_lo;
    a.ring_out_hi = h->stall_test ? s.sync + 8 : s.ring_out_hi;
    a.ring_tickets = s.sync + 4;
    a.ring_timeout = s.sync + 2;
    a.ring_step = h->phase;
    a.nb_hi = s.nblk - (int)(last_row_first_item / h->tpb);
    a.nb_lo = (a.nxv + h->tpb - 1) / h->tpb;
    a.rot = a.nb_hi;
  }
  fill_reduce(h, s, a.red, step_index);
  return a;
}

int launch_accelerate(lbm_lattice* h, Slab& s, int cur)
{
  if (!s.owns_accel_row) return 0;
  const int nx = h->p.nx;
  lbm::accelerate_row_kernel<<<(nx + 255) / 256, 256, 0, s.stream>>>(s.buf[cur], s.flags, s.ps, nx, s.accel_row,
                                                                    h->a1, h->a2);
  CK(cudaGetLastError());
  h->last_launches++;
  return 0;
}

// ghost zones of buffer `cur` <- the neighbours' boundary rows; with a ring: publishes phase + 1
int launch_refresh(lbm_lattice* h, Slab& s, int cur)
{
  const int nx = h->p.nx;
  unsigned* out_lo = h->ring ? (h->stall_test ? s.sync + 8 : s.ring_out_lo) : nullptr;
  unsigned* out_hi = h->ring ? (h->stall_test ? s.sync + 8 : s.ring_out_hi) : nullptr;
  const long long items = 18LL * G * nx;
  const int nb = (int)std::min<long long>((items / 4 + 255) / 256, 148LL * 4);
  if (nx % 4 == 0)
    lbm::ghost_refresh_kernel<<<nb, 256, 0, s.stream>>>(s.buf[cur], s.ps, nx, s.rows, s.gz_lo[cur], s.lo_ps,
                                                       s.gz_hi[cur], s.hi_ps, out_lo, out_hi, s.sync + 4, h->phase);
  else
    lbm::ghost_refresh_scalar_kernel<<<nb, 256, 0, s.stream>>>(s.buf[cur], s.ps, nx, s.rows, s.gz_lo[cur], s.lo_ps,
                                                              s.gz_hi[cur], s.hi_ps, out_lo, out_hi, s.sync + 4,
                                                              h->phase);
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

// all ranks have reached this point and their streams are idle (multi-process mode only); the sum
// of `flag` over the ranks comes back in *flag_sum, so that every rank can fail together
int comm_barrier(lbm_lattice* h, long long flag = 0, long long* flag_sum = nullptr)
{
  if (flag_sum) *flag_sum = flag;
  if (!h->comm) return 0;
  Slab& s = h->slabs[0];
  NcclApi* n = nccl_api();
  long long* d = h->comm->scratch64;
  CK(cudaMemcpyAsync(d, &flag, sizeof flag, cudaMemcpyHostToDevice, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  NK(n->AllReduce(d, d, 1, ncclInt64, ncclSum, h->comm->nccl, s.stream));
  long long out = 0;
  CK(cudaMemcpyAsync(&out, d, sizeof out, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  if (flag_sum) *flag_sum = out;
  return 0;
}

// NCCL flavour of the halo (one-step kernel only): my first row's 4,7,8 go down, my last row's
// 2,5,6 go up, straight from / into the plane rows of buffer `b` (each a contiguous run of nx floats)
int nccl_halo_exchange(lbm_lattice* h, Slab& s, int b)
{
  NcclApi* n = nccl_api();
  Comm* c = h->comm;
  const int nx = h->p.nx;
  const int lo = (c->rank + c->world - 1) % c->world, hi = (c->rank + 1) % c->world;
  float* base = s.buf[b];
  const long long first = (long long)G * nx, last = (long long)(G + s.rows - 1) * nx;
  ncclResult_t r = n->GroupStart();
  for (int i = 0; i < 3 && r == ncclSuccess; i++) {
    r = n->Send(base + LO_PLANES[i] * s.ps + first, nx, ncclFloat, lo, c->nccl, s.stream);
    if (r == ncclSuccess) r = n->Send(base + HI_PLANES[i] * s.ps + last, nx, ncclFloat, hi, c->nccl, s.stream);
    if (r == ncclSuccess) r = n->Recv(base + LO_PLANES[i] * s.ps + last + nx, nx, ncclFloat, hi, c->nccl, s.stream);
    if (r == ncclSuccess) r = n->Recv(base + HI_PLANES[i] * s.ps + first - nx, nx, ncclFloat, lo, c->nccl, s.stream);
  }
  const ncclResult_t e = n->GroupEnd();     // always paired with GroupStart
  if (r == ncclSuccess) r = e;
  if (r != ncclSuccess) return fail("NCCL error during the halo exchange: %s", n->GetErrorString(r));
  return 0;
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
    const StepArgs a = make_args(h, s, c, 1, i, 0);
    CK(launch_step(h->vec, h->tpb, a, s.nblk, s.stream, h->use_pdl && i > 0));
    c ^= 1;
  }
  lbm::reduce_partials_kernel<<<h->chunk, 256, 0, s.stream>>>(s.partials, s.np, s.nblk, s.totals, s.counter);
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

// ---- S timesteps per pass (the streaming kernel) -------------------------------------------------
// Fewer bytes per update (72 / S + halo) at the price of lower occupancy and ~7 % redundant columns:
// wins once a slab streams from HBM and loses when it lives in L2, so "auto" turns it on from
// 8 M cells (0.6 GB of state) per slab up.
int stream_steps(const lbm_lattice* h)
{
  // decided from global quantities only: every rank of a ring must take the same path
  const int min_rows = h->p.ny / h->world;
  const int nx = h->p.nx;
  if (h->fuse_mode == 1 || nx % 16 != 0 || nx < 128 || min_rows < 4 * G) return 1;
  if (h->comm && h->comm->halo != HALO_P2P) return 1;
  if (!encode_tiled_fn()) return 1;
  const int s = STREAM_CFGS[h->stream_cfg].s;
  if (h->fuse_mode >= 2) return s;
  return ((long long)min_rows * nx >= (8LL << 20)) ? s : 1;
}

template <int S, int NW, int K0, int MINB>
int launch_stream_t(const Slab& sl, const StreamArgs& a, const lbm::StepReduce& r, int ntiles, cudaStream_t st, bool pdl)
{
  constexpr int SMEM = lbm::stream_smem_bytes(S, NW, K0);
  // programmatic dependent launch: the next pass's blocks are scheduled while this pass's last tiles
  // are still running and park at griddepcontrol.wait (kernel prologue) until it has completed
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)ntiles);
  cfg.blockDim = dim3((S * NW + 1) * 32);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, lbm::lbm_stream_kernel<S, NW, K0, MINB>, sl.tm_state, sl.tm_flags, sl.tm_state_w,
                        sl.tm_flags_w, a, r));
  return 0;
}

template <int S, int NW, int K0, int MINB>
cudaError_t configure_stream_t()
{
  return cudaFuncSetAttribute(lbm::lbm_stream_kernel<S, NW, K0, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              lbm::stream_smem_bytes(S, NW, K0));
}

#define LBM_STREAM_DISPATCH(idx, CALL)                 \
  switch (idx) {                                       \
    case 1: return CALL(3, 6, 3, 1);                   \
    case 2: return CALL(2, 8, 3, 1);                   \
    case 3: return CALL(4, 4, 2, 1);                   \
    case 4: return CALL(2, 3, 2, 3);                   \
    case 5: return CALL(2, 4, 2, 2);                   \
    case 6: return CALL(3, 4, 2, 1);                   \
    default: return CALL(2, 4, 3, 2);                  \
  }

// the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
cudaError_t configure_stream(int cfg)
{
#define LBM_CALL(S, NW, K0, MINB) configure_stream_t<S, NW, K0, MINB>()
  LBM_STREAM_DISPATCH(cfg, LBM_CALL)
#undef LBM_CALL
}

// t -> t+S on every owned row of buffer cur^1 (+ the neighbours' ghost zones); slots: S steps
int launch_stream(lbm_lattice* h, Slab& s, int cur, int fuse_last, int slot, long long step_index, bool pdl)
{
  StreamArgs a{};
  a.dst = s.buf[cur ^ 1];
  a.flags = s.flags;
  a.ps = s.ps;
  a.nx = h->p.nx;
  a.rows = s.rows;
  a.tiles_x = s.tiles_x;
  a.tiles_y = s.tiles_y;
  a.tile_h = s.tile_h;
  a.tall_rows = s.tall_rows;
  a.tile_h2 = s.tile_h2;
  a.src_plane0 = 9 * cur;
  a.omega = h->p.omega;
  a.a1 = h->a1;
  a.a2 = h->a2;
  a.fuse_last = fuse_last;
  a.ghost_lo = s.gz_lo[cur ^ 1];
  a.ghost_hi = s.gz_hi[cur ^ 1];
  a.ps_lo = s.lo_ps;
  a.ps_hi = s.hi_ps;
  a.partials = s.partials + (long long)slot * s.np;
  a.np = s.np;
  if (h->ring) {
    a.ring_in = s.sync;
    a.ring_out_lo = h->stall_test ? s.sync + 8 : s.ring_out_lo;
    a.ring_out_hi = h->stall_test ? s.sync + 8 : s.ring_out_hi;
    a.ring_tickets = s.sync + 4;
    a.ring_timeout = s.sync + 2;
    a.ring_phase = h->phase;
    int n_lo = 0, n_hi = 0;      // the kernel's own conditions, counted over the rows of tiles
    for (int by = 0; by < s.tiles_y; by++) {
      n_lo += tile_row0(s, by) < 2 * G;
      n_hi += tile_row1(s, by) > s.rows;
    }
    a.ring_n_lo = n_lo * s.tiles_x;
    a.ring_n_hi = n_hi * s.tiles_x;
  }
  lbm::StepReduce r;
  fill_reduce(h, s, r, step_index);
  const int nt = s.tiles_x * s.tiles_y;
  a.trace = s.trace;
#define LBM_CALL(S, NW, K0, MINB) launch_stream_t<S, NW, K0, MINB>(s, a, r, nt, s.stream, pdl)
  LBM_STREAM_DISPATCH(h->stream_cfg, LBM_CALL)
#undef LBM_CALL
}

// one-step launches of one timestep on every slab (+ the ordering / halo launches of the slow variants)
int launch_one_step(lbm_lattice* h, int cur, int fuse, int slot, long long step_index, bool pdl)
{
  Comm* comm = h->comm;
  const bool launches_order = h->ring && !h->ring_in_kernel;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    // my ring neighbours must have finished the previous phase: their stores into my ghost rows are
    // complete and they no longer read the ghost rows I am about to overwrite
    if (launches_order) {
      lbm::wait_neighbours_kernel<<<1, 2, 0, s.stream>>>(s.sync, h->phase, s.sync + 2);
      CK(cudaGetLastError());
      h->last_launches++;
    }
    const StepArgs a = make_args(h, s, cur, fuse, slot, step_index);
    CK(launch_step(h->vec, h->tpb, a, s.nblk, s.stream, pdl));
    h->last_launches++;
    if (launches_order) {
      lbm::signal_neighbours_kernel<<<1, 2, 0, s.stream>>>(h->stall_test ? s.sync + 8 : s.ring_out_lo,
                                                           h->stall_test ? s.sync + 8 : s.ring_out_hi, h->phase + 1);
      CK(cudaGetLastError());
      h->last_launches++;
    } else if (comm && comm->halo == HALO_NCCL) {
      if (nccl_halo_exchange(h, s, cur ^ 1)) return 1;
      h->last_launches++;
    }
  }
  h->phase++;
  return 0;
}

void set_config_string(lbm_lattice* h);

// ---- the persistent small-lattice kernel (lbm_resident.cuh) ---------------------------------------
// `n` timesteps from buffer `cur` into buffer cur^1 in one cooperative launch; partial sums of step i
// go to slot i
int launch_resident(lbm_lattice* h, Slab& s, int cur, int n, int fuse_after)
{
  lbm::ResidentArgs a{};
  a.src = s.buf[cur];
  a.dst = s.buf[cur ^ 1];
  a.flags = s.flags;
  a.ps = s.ps;
  a.nx = h->p.nx;
  a.ny = s.rows;
  a.R = h->res_rows;
  a.nsteps = n;
  a.fuse_after = fuse_after;
  a.np = s.np;
  a.omega = h->p.omega;
  a.a1 = h->a1;
  a.a2 = h->a2;
  a.partials = s.partials;
  a.inbox = s.inbox;
  a.base = h->res_base;
  a.timed_out = s.sync + 2;
  a.stall_block = env_int("LBM_TEST_RESIDENT_STALL", -1);
#ifdef LBM_RES_TRACE
  static long long* trace = nullptr;
  if (!trace) { CK(cudaMallocManaged(&trace, sizeof(long long) * 2 * 16 * 6)); memset(trace, 0, sizeof(long long) * 192); }
  a.trace = trace;
  if (const char* path = getenv("LBM_RES_TRACE_FILE")) {      // dump what the PREVIOUS launch recorded
    cudaStreamSynchronize(s.stream);
    if (FILE* fp = fopen(path, "w")) {
      for (int th = 0; th < 2; th++)
        for (int i = 0; i < 16; i++) {
          const long long* r = trace + (th * 16 + i) * 6;
          fprintf(fp, "thread %s step %d: compute+send %lld  poll %lld  barrier %lld  reduce %lld | step %lld\n", th ? "last" : "0", 64 + i,
                  r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], i ? r[0] - (r - 6)[0] : 0LL);
        }
      fclose(fp);
    }
  }
#endif
  h->res_base += (unsigned)n;
  h->last_launches++;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)h->res_nblk);
  cfg.blockDim = dim3((unsigned)h->res_tpb);
  cfg.dynamicSmemBytes = lbm::resident_smem_bytes(h->res_rows, h->p.nx, h->res_tpb);
  cfg.stream = s.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all blocks co-resident, or the launch fails
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = env_int("LBM_TEST_RESIDENT_LAUNCH_FAIL", 0) ? cudaErrorCooperativeLaunchTooLarge : cudaSuccess;
  if (e == cudaSuccess) {
    switch (h->res_cpt) {
      case 1: e = cudaLaunchKernelEx(&cfg, lbm::lbm_resident_kernel<1>, a); break;
      case 2: e = cudaLaunchKernelEx(&cfg, lbm::lbm_resident_kernel<2>, a); break;
      case 4: e = cudaLaunchKernelEx(&cfg, lbm::lbm_resident_kernel<4>, a); break;
      default: e = cudaLaunchKernelEx(&cfg, lbm::lbm_resident_kernel<8>, a); break;
    }
  }
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources || e == cudaErrorNotSupported) {
    // the device cannot hold the whole grid right now (e.g. an MPS partition smaller than the occupancy
    // query assumed): nothing ran; the caller carries on with one launch per timestep
    cudaGetLastError();
    h->last_launches--;
    return 2;
  }
  if (e != cudaSuccess) return fail("CUDA error launching the persistent kernel: %s", cudaGetErrorString(e));
  return 0;
}

// shared-memory opt-in and co-residency of the instantiation for `cpt` cells per thread
int resident_prepare(int cpt, int tpb, size_t smem, int* blocks_per_sm)
{
  const void* fn = cpt == 1 ? (const void*)lbm::lbm_resident_kernel<1>
                   : cpt == 2 ? (const void*)lbm::lbm_resident_kernel<2>
                   : cpt == 4 ? (const void*)lbm::lbm_resident_kernel<4>
                              : (const void*)lbm::lbm_resident_kernel<8>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, tpb, smem));
  return 0;
}

// second stage of the reduction for the `nsteps` steps just launched; their partials are per tile
// (streaming passes) or per block (one-step launches: count = 0; persistent kernel: its block count)
int reduce_chunk(lbm_lattice* h, int nsteps, bool tiles, int count = 0)
{
  if (h->reduce_per_step) return 0;     // done by the last block of every launch
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    lbm::reduce_partials_kernel<<<nsteps, 256, 0, s.stream>>>(s.partials, s.np,
                                                             count ? count : tiles ? s.tiles_x * s.tiles_y : s.nblk,
                                                             s.totals, s.counter);
    CK(cudaGetLastError());
    lbm::advance_counter_kernel<<<1, 1, 0, s.stream>>>(s.counter, nsteps);
    CK(cudaGetLastError());
  }
  h->last_launches += 2 * (long long)h->slabs.size();
  return 0;
}

// the time loop proper: `iters` timesteps starting at absolute step index `step0` of this run
int run_steps(lbm_lattice* h, int iters, long long step0, bool last_segment)
{
  const size_t nslab = h->slabs.size();
  int remaining = iters;
  int cur = h->cur;
  long long step = step0;
  // steps still to come after this segment keep the pre-acceleration going
  const int more = last_segment ? 0 : 1;

  const int S = stream_steps(h);
  if (S > 1) {
    while (remaining >= S) {
      const int passes = std::min(remaining / S, std::max(1, h->chunk / S));
      for (int j = 0; j < passes; j++) {
        const int fuse_last = (remaining - S * (j + 1) + more) > 0;
        for (auto& s : h->slabs) {
          CK(cudaSetDevice(s.device));
          if (launch_stream(h, s, cur, fuse_last, S * j, step, h->stream_pdl && j > 0)) return 1;
        }
        h->phase++;
        h->last_launches += (long long)nslab;
        step += S;
        cur ^= 1;
      }
      if (reduce_chunk(h, S * passes, true)) return 1;
      remaining -= S * passes;
    }
  }

  // small lattice on one GPU: the remaining steps as persistent launches of up to res_chunk steps (a
  // single step -- lbm_step -- stays a plain launch)
  if (h->resident && remaining >= 2) {
    Slab& s = h->slabs[0];
    CK(cudaSetDevice(s.device));
    while (remaining > 0) {
      const int n = std::min(remaining, h->res_chunk);
      const int rc = launch_resident(h, s, cur, n, (remaining - n + more) > 0);
      if (rc == 2) {                  // not launchable here: this handle uses the one-step launches from now on
        h->resident = false;
        set_config_string(h);
        break;
      }
      if (rc) return 1;
      if (reduce_chunk(h, n, false, h->res_nblk)) return 1;
      cur ^= 1;                       // a launch reads one buffer and leaves its result in the other
      remaining -= n;
      step += n;
    }
  }

  if (nslab == 1 && h->world == 1 && h->use_graph) {
    Slab& s = h->slabs[0];
    while (remaining > h->chunk) {
      if (!s.graph[cur] && build_graph(h, s, cur)) return 1;
      CK(cudaGraphLaunch(s.graph[cur], s.stream));
      h->last_launches += h->chunk + 2;
      remaining -= h->chunk;                    // chunk is even: parity unchanged
      step += h->chunk;
    }
  }

  // remaining steps as plain launches, reduced chunk by chunk; the very last step of the run
  // leaves the state un-accelerated
  while (remaining > 0) {
    const int n = std::min(remaining, h->chunk);
    for (int i = 0; i < n; i++, step++) {
      const int fuse = (remaining - i - 1 + more) > 0;
      if (launch_one_step(h, cur, fuse, i, step, h->use_pdl && h->world == 1 && i > 0)) return 1;
      cur ^= 1;
    }
    if (reduce_chunk(h, n, false)) return 1;
    remaining -= n;
  }
  h->cur = cur;
  return 0;
}

int run_impl(lbm_lattice* h, int iters, double* av_out)
{
  h->last_ms = 0;
  h->last_launches = 0;
  if (iters < 0) return fail("lbm_run: negative iteration count");
  if (iters == 0) return 0;
  if (h->poisoned) return fail("lbm_run: an earlier run on this handle failed half-way; create a new one");
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
    if (comm && comm->halo == HALO_NCCL) {
      if (nccl_halo_exchange(h, s, h->cur)) return 1;
    } else if (launch_refresh(h, s, h->cur)) {
      return 1;
    }
  }
  if (h->ring) h->phase++;

  std::fill(av_out, av_out + iters, 0.0);
  std::vector<double> tmp;
  int rc = 0;
  if (!h->reduce_per_step) {
    rc = run_steps(h, iters, 0, true);
  } else {
    // LBM_REDUCE=step: every launch pushes its steps' slab totals to all ranks; the arrays hold
    // ALLRED_CAP steps, so longer runs are drained segment by segment
    tmp.resize((size_t)ALLRED_CAP * h->world);
    for (long long done = 0; done < iters && rc == 0;) {
      const int seg = (int)std::min<long long>(iters - done, ALLRED_CAP);
      rc = run_steps(h, seg, done, done + seg == iters);
      if (rc) break;
      if (sync_all(h) || comm_barrier(h)) { rc = 1; break; }      // everybody's pushes have landed
      Slab& s = h->slabs[0];      // every slab holds the same numbers
      CK(cudaSetDevice(s.device));
      CK(cudaMemcpy(tmp.data(), s.allred, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
      for (int r = 0; r < h->world; r++)       // fixed rank order: deterministic, identical everywhere
        for (int t = 0; t < seg; t++) av_out[done + t] += tmp[(size_t)r * ALLRED_CAP + (done + t) % ALLRED_CAP];
      done += seg;
    }
  }
  if (rc) { h->poisoned = h->ring; return 1; }

  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    CK(cudaEventRecord(s.ev_end, s.stream));
  }
  if (sync_all(h)) { h->poisoned = h->ring; return 1; }
  float ms_max = 0;
  for (auto& s : h->slabs) {
    float ms = 0;
    CK(cudaSetDevice(s.device));
    CK(cudaEventElapsedTime(&ms, s.ev_begin, s.ev_end));
    ms_max = std::max(ms_max, ms);
  }
  h->last_ms = ms_max;
  if (const char* path = getenv("LBM_STREAM_TRACE")) {     // tuning aid: the latest pass's tile schedule
    Slab& s = h->slabs[0];
    if (s.trace) {
      const size_t nt = (size_t)s.tiles_x * s.tiles_y;
      std::vector<unsigned long long> t(4 * nt);
      CK(cudaMemcpy(t.data(), s.trace, sizeof(unsigned long long) * 4 * nt, cudaMemcpyDeviceToHost));
      if (FILE* fp = fopen(path, "w")) {
        fprintf(fp, "tile,bx,by,sm,block,start_ns,end_ns\n");
        for (size_t i = 0; i < nt; i++)
          fprintf(fp, "%zu,%zu,%zu,%llu,%llu,%llu,%llu\n", i, i % s.tiles_x, i / s.tiles_x, t[4 * i], t[4 * i + 2],
                  t[4 * i + 1], t[4 * i + 3]);
        fclose(fp);
      }
    }
  }

  // a neighbour that never showed up: every slab of the ring fails together
  long long timed_out = 0;
  if (h->ring) {
    for (auto& s : h->slabs) {
      unsigned t = 0;
      CK(cudaSetDevice(s.device));
      CK(cudaMemcpy(&t, s.sync + 2, sizeof t, cudaMemcpyDeviceToHost));
      timed_out += t;
    }
    long long all = timed_out;
    if (comm_barrier(h, timed_out, &all)) { h->poisoned = true; return 1; }
    if (all) {
      h->poisoned = true;
      return fail("lbm_run: timed out waiting for a neighbour GPU (slab %d of %d)", h->slabs[0].rank, h->world);
    }
  }

  if (h->resident) {
    unsigned t = 0;
    Slab& s = h->slabs[0];
    CK(cudaSetDevice(s.device));
    CK(cudaMemcpy(&t, s.sync + 2, sizeof t, cudaMemcpyDeviceToHost));
    if (t) {
      h->poisoned = true;
      return fail("lbm_run: the persistent kernel timed out waiting for a neighbouring block's halo");
    }
  }

  if (!h->reduce_per_step) {
    tmp.resize((size_t)iters);
    if (!comm) {
      for (auto& s : h->slabs) {          // fixed slab order: the cross-GPU sum is deterministic
        CK(cudaSetDevice(s.device));
        CK(cudaMemcpy(tmp.data(), s.totals, sizeof(double) * iters, cudaMemcpyDeviceToHost));
        for (int t = 0; t < iters; t++) av_out[t] += tmp[t];
      }
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
  h->chunk = std::max(12, env_int("LBM_CHUNK", 120)) / 12 * 12;     // a multiple of every S
  h->use_graph = env_int("LBM_GRAPH", 1) != 0;
  h->use_pdl = env_int("LBM_PDL", -1) != 0;   // -1 = decide per slab size (create_slab)
  h->pad = std::max(0, env_int("LBM_PLANE_PAD", 0));
  h->fuse_mode = env_int("LBM_FUSE", -1);
  // LBM_FUSE=S picks the default tile shape for S steps per pass; LBM_STREAM_CFG picks any shape
  int cfg = 0;
  if (h->fuse_mode == 3) cfg = 1;
  if (h->fuse_mode == 4) cfg = 3;
  cfg = env_int("LBM_STREAM_CFG", cfg);
  h->stream_cfg = std::min(N_STREAM_CFGS - 1, std::max(0, cfg));
  if (h->fuse_mode >= 2) h->fuse_mode = STREAM_CFGS[h->stream_cfg].s;
  h->tile_h_max = std::max(0, env_int("LBM_TILE_H", 0));
  h->stream_pdl = env_int("LBM_STREAM_PDL", 1) != 0;
  h->reduce_per_step = env_is("LBM_REDUCE", "step") && h->world > 1 && h->world <= lbm::MAX_REDUCE_WORLD;
  h->stall_test = env_int("LBM_TEST_RING_STALL", -1) >= 0;
}

// Tile heights of the streaming kernel.  A tile recomputes S-1 rows of each vertical neighbour and
// pays a pipeline fill, so tiles should be tall; but every tile costs the same and the hardware hands
// them out in index order, so with tall tiles only the end of a pass leaves SMs idle for up to a whole
// tile (6 % of a pass on a 16384 x 2048 slab).  Hence two zones: tall tiles (tile_h) first, and the
// last ~1.5 waves' worth of rows in short tiles (tile_h2), which fill the tail.  All heights are
// whole batches of NW rows.  Decided from the ring's smallest slab, so every slab uses the same heights.
void choose_tiles(const lbm_lattice* h, Slab& s, int min_rows)
{
  const StreamCfg& c = STREAM_CFGS[h->stream_cfg];
  const int nx = h->p.nx;
  s.tiles_x = (nx + lbm::S_OUT_W - 1) / lbm::S_OUT_W;
  const int extra = 2 * (c.s - 1);
  auto whole_batches = [&](int hh) { return std::max(((hh + extra + c.nw - 1) / c.nw) * c.nw - extra, 1); };
  if (h->tile_h_max > 0) {                        // LBM_TILE_H: one height everywhere
    s.tile_h = s.tile_h2 = whole_batches(h->tile_h_max);
    s.tall_rows = 0;
  } else {
    s.tile_h = whole_batches(32 * c.nw - extra - c.nw + 1);      // 126 rows for S = 2, NW = 4
    s.tile_h2 = whole_batches(std::max(8 * c.nw - extra - c.nw + 1, 8));
    const int slots = 148 * c.minb;
    const int reserve = (3 * slots / (2 * s.tiles_x) + 1) * s.tile_h2;     // ~1.5 waves of short tiles
    s.tall_rows = std::max(0, (min_rows - reserve) / s.tile_h);
  }
  const int rest = s.rows - s.tall_rows * s.tile_h;
  s.tiles_y = s.tall_rows + (std::max(rest, 0) + s.tile_h2 - 1) / s.tile_h2;
}

int make_tensor_maps(const lbm_lattice* h, Slab& s)
{
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from this driver");
  const StreamCfg& c = STREAM_CFGS[h->stream_cfg];
  const cuuint64_t nx = (cuuint64_t)h->p.nx, nrows = (cuuint64_t)(s.rows + 2 * G);
  for (int wrap = 0; wrap < 2; wrap++) {
    {
      const cuuint64_t dims[3] = {nx, nrows, 18};
      const cuuint64_t strides[2] = {nx * 4, (cuuint64_t)s.ps * 4};
      const cuuint32_t box[3] = {(cuuint32_t)(wrap ? 4 : lbm::S_TILE_W), (cuuint32_t)c.nw, 1};
      const cuuint32_t es[3] = {1, 1, 1};
      const CUresult r = enc(wrap ? &s.tm_state_w : &s.tm_state, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s.base, dims,
                             strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (populations) failed with CUresult %d", (int)r);
    }
    {
      const cuuint64_t dims[2] = {nx, nrows};
      const cuuint64_t strides[1] = {nx};
      const cuuint32_t box[2] = {(cuuint32_t)(wrap ? 16 : lbm::S_FLAG_BOX), (cuuint32_t)c.nw};
      const cuuint32_t es[2] = {1, 1};
      const CUresult r = enc(wrap ? &s.tm_flags_w : &s.tm_flags, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, s.flags, dims,
                             strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (flags) failed with CUresult %d", (int)r);
    }
  }
  return 0;
}

// Small lattice on one GPU: can the time loop be the persistent kernel?  Block b keeps R whole rows in
// shared memory (two copies) and all blocks must be resident at once -- one block per SM, so
// R = ceil(ny / SMs) -- which bounds the lattice at roughly (9R + 6) * nx * 8 bytes <= 227 KB per block:
// 512 x 512 fits, 1024 x 1024 does not (and does not need it: its steps are 10 us of real work).
int decide_resident(lbm_lattice* h, Slab& s)
{
  const int mode = env_int("LBM_RESIDENT", -1);       // 0 never, 1 whenever it fits, -1 auto
  if (mode == 0) return 0;
  const int nx = h->p.nx, ny = s.rows;
  if (mode < 0 && (long long)nx * ny > (long long)env_int("LBM_RESIDENT_MAX_CELLS", 1 << 19)) return 0;
  int sms = 0, coop = 0, smem_max = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device));
  CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s.device));
  CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, s.device));
  if (!coop || sms < 1) return 0;
  int R = std::max((ny + sms - 1) / sms, env_int("LBM_RES_ROWS", 1));
  R = std::min(R, ny);
  int tpb = std::min(512, std::max(32, env_int("LBM_RES_TPB", 512)));
  tpb = std::min(tpb, (R * nx + 31) / 32 * 32) / 32 * 32;
  const size_t smem = lbm::resident_smem_bytes(R, nx, tpb);
  if (smem > (size_t)smem_max) return 0;
  const int nblk = (ny + R - 1) / R;
  int cpt = (R * nx + tpb - 1) / tpb;                 // cells per thread: instantiated for 1, 2, 4, 8
  if (cpt > lbm::RES_MAX_CPT) return 0;
  cpt = cpt <= 1 ? 1 : cpt <= 2 ? 2 : cpt <= 4 ? 4 : 8;
  int per_sm = 0;
  if (resident_prepare(cpt, tpb, smem, &per_sm)) return 1;
  if (nblk > per_sm * sms) return 0;
  h->res_cpt = cpt;
  h->resident = true;
  h->res_rows = R;
  h->res_tpb = tpb;
  h->res_nblk = nblk;
  h->res_chunk = std::max(2, env_int("LBM_RES_CHUNK", 1200));
  const size_t words = (size_t)nblk * 2 * 2 * 3 * nx;
  CK(cudaMalloc(&s.inbox, sizeof(unsigned long long) * words));
  CK(cudaMemsetAsync(s.inbox, 0, sizeof(unsigned long long) * words, s.stream));
  return 0;
}

// device objects of one slab.  obstacles: `rows_given` rows starting at lattice row `row_first`
// (the whole grid in one-process mode, the slab's own rows in rank mode); flag rows whose lattice
// row is not among them (ghost rows in rank mode) are left for the neighbours to push.
int create_slab(lbm_lattice* h, Slab& s, const int* obstacles, int row_first, int rows_given,
                long long* fluid_cells)
{
  const int nx = h->p.nx, ny = h->p.ny;
  const long long cells = (long long)(s.rows + 2 * G) * nx;
  s.ps = plane_stride(s.rows, nx, h->pad);
  s.nvec = (long long)s.rows * (nx / h->vec);
  if (s.nvec >= (1LL << 31)) return fail("lbm_create: slab too large for 32-bit work index");
  s.nblk = (int)((s.nvec + h->tpb - 1) / h->tpb);
  // PDL pays once a step is longer than a launch: measured +2 % at 1024^2 (2048 blocks), +10 % at
  // 128^2 but -50 % at 256^2 inside graphs (profiles/r1_tuning.md), so auto = multi-wave grids only
  if (env_int("LBM_PDL", -1) < 0) h->use_pdl = s.nblk >= 148 * 8;
  s.owns_accel_row = (ny - 2 >= s.y0 && ny - 2 < s.y0 + s.rows);
  s.accel_row = ny - 2 - s.y0 + G;

  CK(cudaSetDevice(s.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, s.device));
  if (prop.major < 10)
    return fail("device %d (%s) is not an sm_100-class GPU; this library is built for sm_100a only",
                s.device, prop.name);
  CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&s.ev_begin));
  CK(cudaEventCreate(&s.ev_end));
  const size_t bytes = slab_bytes_total(s.ps, s.rows, nx, h->world);
  CK(cudaMalloc(&s.base, bytes));
  CK(cudaMemsetAsync(s.base, 0, bytes, s.stream));
  s.buf[0] = buf_of(s.base, s.ps, 0);
  s.buf[1] = buf_of(s.base, s.ps, 1);
  s.sync = sync_of(s.base, s.ps);
  s.flags = flags_of(s.base, s.ps);
  s.allred = allred_of(s.base, s.ps, s.rows, nx);
  const int min_rows = ny / h->world;
  if (stream_steps(h) > 1) {
    choose_tiles(h, s, min_rows);
    if (make_tensor_maps(h, s)) return 1;
    const cudaError_t e = configure_stream(h->stream_cfg);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute(shared memory): %s", cudaGetErrorString(e));
  }
  s.np = std::max(s.nblk, s.tiles_x * s.tiles_y);
  if (h->world == 1 && !h->comm && stream_steps(h) == 1 && decide_resident(h, s)) return 1;
  if (h->resident) s.np = std::max(s.np, h->res_nblk);
  if (getenv("LBM_STREAM_TRACE") && s.tiles_x * s.tiles_y > 0) {
    CK(cudaMalloc(&s.trace, sizeof(unsigned long long) * 4 * (size_t)s.tiles_x * s.tiles_y));
    CK(cudaMemsetAsync(s.trace, 0, sizeof(unsigned long long) * 4 * (size_t)s.tiles_x * s.tiles_y, s.stream));
  }
  CK(cudaMalloc(&s.partials, sizeof(double) * (size_t)std::max(h->chunk, h->resident ? h->res_chunk : 0) * s.np));
  CK(cudaMalloc(&s.counter, sizeof(long long)));
  CK(cudaMemsetAsync(s.counter, 0, sizeof(long long), s.stream));

  // flags: bit 0 obstacle, bit 1 fluid cell of the accelerated row (global ny-2); ghost rows carry
  // the flags of the lattice rows they mirror (periodic in y)
  std::vector<uint8_t> fl((size_t)cells, 0);
  long long fluid = 0;
#pragma omp parallel for reduction(+ : fluid) schedule(static)
  for (int r = 0; r < s.rows + 2 * G; r++) {
    const int gy = ((s.y0 + r - G) % ny + ny) % ny;
    int src = gy - row_first;
    if (src < 0 || src >= rows_given) continue;
    const bool owned = r >= G && r < G + s.rows;
    const int* orow = obstacles + (size_t)src * nx;
    uint8_t* frow = fl.data() + (size_t)r * nx;
    for (int x = 0; x < nx; x++) {
      const bool ob = orow[x] != 0;
      frow[x] = (uint8_t)((ob ? lbm::FLAG_OBSTACLE : 0) | ((!ob && gy == ny - 2) ? lbm::FLAG_ACCEL : 0));
      if (owned) fluid += !ob;
    }
  }
  *fluid_cells += fluid;
  CK(cudaMemcpyAsync(s.flags, fl.data(), (size_t)cells, cudaMemcpyHostToDevice, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return 0;
}

// ghost destinations of slab `s` inside the allocations `lo_base` / `hi_base` of its ring
// neighbours, whose geometry is (lo_rows, lo_ps) / hi_ps
void wire_ghosts(Slab& s, int nx, char* lo_base, int lo_rows, long long lo_ps, char* hi_base,
                 long long hi_ps)
{
  s.lo_base = lo_base;
  s.hi_base = hi_base;
  s.lo_rows = lo_rows;
  s.lo_ps = lo_ps;
  s.hi_ps = hi_ps;
  for (int b = 0; b < 2; b++) {
    float* lo_buf = buf_of(lo_base, lo_ps, b);
    float* hi_buf = buf_of(hi_base, hi_ps, b);
    s.gz_lo[b] = lo_buf + (long long)(G + lo_rows) * nx;
    s.gz_hi[b] = hi_buf;
    for (int i = 0; i < 3; i++) {
      s.ghost_lo[b][i] = lo_buf + LO_PLANES[i] * lo_ps + (long long)(G + lo_rows) * nx;
      s.ghost_hi[b][i] = hi_buf + HI_PLANES[i] * hi_ps + (long long)(G - 1) * nx;
    }
  }
  // my lower neighbour counts me as its UPPER neighbour (its sync[1]); the upper one as its LOWER
  s.ring_out_lo = sync_of(lo_base, lo_ps) + 1;
  s.ring_out_hi = sync_of(hi_base, hi_ps) + 0;
}

void set_config_string(lbm_lattice* h)
{
  char cfg[520], stream[200] = "";
  const int S = stream_steps(h);
  const bool multi = h->world > 1;
  const char* red = (multi && h->reduce_per_step) ? " reduce=in-kernel-allreduce-per-step" : "";
  const char* mode = !multi ? "single-gpu"
                     : (h->comm && h->comm->halo == HALO_NCCL) ? "ranks+nccl-sendrecv"
                     : h->comm ? ((S > 1 || h->ring_in_kernel) ? "ranks+ipc-peer-stores+in-kernel-ring"
                                                              : "ranks+ipc-peer-stores+wait/signal-kernels")
                               : ((S > 1 || h->ring_in_kernel) ? "one-process+peer-stores+in-kernel-ring"
                                                              : "one-process+peer-stores+wait/signal-kernels");
  if (S > 1) {
    const StreamCfg& c = STREAM_CFGS[h->stream_cfg];
    snprintf(stream, sizeof stream, " stream=tma(S=%d,nw=%d,stages=%d,blocks/sm=%d,pdl=%d) tile=%dx%d*%d+%dx%d*%d",
             c.s, c.nw, c.k0, c.minb, (int)h->stream_pdl, lbm::S_OUT_W, h->slabs[0].tile_h, h->slabs[0].tall_rows,
             lbm::S_OUT_W, h->slabs[0].tile_h2, h->slabs[0].tiles_y - h->slabs[0].tall_rows);
  }
  if (h->resident)
    snprintf(stream, sizeof stream, " resident=smem(rows/block=%d,tpb=%d,blocks=%d,smem=%zuKB,steps/launch=%d)",
             h->res_rows, h->res_tpb, h->res_nblk, lbm::resident_smem_bytes(h->res_rows, h->p.nx, h->res_tpb) >> 10,
             h->res_chunk);
  // graph replay and PDL apply to the one-step kernel on one GPU only
  const bool one_step_only = S == 1 && !multi && !h->resident;
  snprintf(cfg, sizeof cfg, "vec=%d tpb=%d chunk=%d graph=%d pdl=%d fuse=%d%s slabs=%d halo=%s%s plane_stride=%lld",
           h->vec, h->tpb, h->chunk, (int)(h->use_graph && one_step_only), (int)(h->use_pdl && one_step_only),
           S, stream, h->world, mode, red, h->slabs[0].ps);
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

void set_ring_mode(lbm_lattice* h)
{
  h->ring = h->world > 1 && !(h->comm && h->comm->halo == HALO_NCCL);
  // Ring ordering inside the ONE-STEP kernel needs every row to start on its own 128-byte line (an
  // interior block must not pull a stale copy of a ghost row's tail into L1 before the boundary
  // block has seen the neighbour's flag); otherwise two tiny wait/signal launches bracket each step.
  // (The streaming kernel reads through TMA, which bypasses L1, and always orders itself.)
  h->ring_in_kernel = h->ring && (h->p.nx % 32 == 0) && !env_is("LBM_RING", "kernels");
}

int create_impl(lbm_lattice** out, const lbm_params* p, const int* obstacles, int first_device,
                int nslab)
{
  if (common_checks(out, p, obstacles)) return 1;
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (nslab < 1 || first_device < 0 || first_device + nslab > ndev)
    return fail("lbm_create: %d GPU(s) requested from device %d but %d visible", nslab, first_device, ndev);
  if (nslab > 1 && p->ny / nslab < G)
    return fail("lbm_create: %d rows cannot be split into %d slabs of >= %d rows", p->ny, nslab, G);
  if (nslab > MAX_WORLD) return fail("lbm_create: at most %d slabs", MAX_WORLD);

  lbm_lattice* h = new_lattice(p, nslab);
  set_ring_mode(h);
  const int nx = p->nx;
  h->slabs.resize(nslab);
  for (int k = 0; k < nslab; k++) {
    Slab& s = h->slabs[k];
    s.device = first_device + k;
    s.rank = k;
    lbm_slab_rows(p->ny, nslab, k, &s.y0, &s.rows);
    if (create_slab(h, s, obstacles, 0, p->ny, &h->tot_cells)) { lbm_destroy(h); return 1; }
  }
  for (int k = 0; k < nslab; k++) {
    Slab& s = h->slabs[k];
    Slab& lo = h->slabs[(k + nslab - 1) % nslab];
    Slab& hi = h->slabs[(k + 1) % nslab];
    wire_ghosts(s, nx, lo.base, lo.rows, lo.ps, hi.base, hi.ps);
    for (int r = 0; r < nslab; r++) s.allred_peer[r] = h->slabs[r].allred;
    if (nslab > 1) {
      cudaSetDevice(s.device);
      for (int r = 0; r < nslab; r++) {
        const Slab* nb = &h->slabs[r];
        const bool neighbour = nb == &lo || nb == &hi;
        if (nb->device == s.device || !(neighbour || h->reduce_per_step)) continue;
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

// one process per GPU: NCCL communicator, global cell count, IPC mapping of the other ranks
int attach_comm(lbm_lattice* h, int rank, int world, const void* unique_id)
{
  NcclApi* n = nccl_api();
  if (!n) return fail("lbm_create_rank: libnccl.so.2 could not be loaded (%s)", dlerror());
  if (!unique_id) return fail("lbm_create_rank: world > 1 needs an ncclUniqueId");
  Slab& s = h->slabs[0];
  Comm* c = h->comm;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof id);
  NK(n->CommInitRank(&c->nccl, world, id, rank));
  CK(cudaMalloc(&c->scratch64, sizeof(long long) * 2));
  CK(cudaMemsetAsync(c->scratch64, 0, sizeof(long long) * 2, s.stream));

  // global number of fluid cells (d2q9-bgk.c:146-152 counts them over the whole grid)
  if (comm_barrier(h, h->tot_cells, &h->tot_cells)) return 1;

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
  CK(cudaMemcpyAsync(d_all + hb * world, &mine, hb, cudaMemcpyHostToDevice, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  NK(n->AllGather(d_all + hb * world, d_all, hb, ncclChar, c->nccl, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  std::vector<cudaIpcMemHandle_t> all(world);
  CK(cudaMemcpy(all.data(), d_all, hb * world, cudaMemcpyDeviceToHost));
  CK(cudaFree(d_all));
  c->peers.assign(world, nullptr);
  for (int r = 0; r < world; r++) {
    if (r == rank || !(r == lo || r == hi || h->reduce_per_step)) continue;
    CK(cudaIpcOpenMemHandle(&c->peers[r], all[r], cudaIpcMemLazyEnablePeerAccess));
  }
  wire_ghosts(s, nx, (char*)c->peers[lo], lo_rows, lo_ps, (char*)c->peers[hi], hi_ps);
  if (h->reduce_per_step) {
    for (int r = 0; r < world; r++) {
      int ry0, rrows;
      lbm_slab_rows(h->p.ny, world, r, &ry0, &rrows);
      s.allred_peer[r] = r == rank ? s.allred : allred_of((char*)c->peers[r], plane_stride(rrows, nx, h->pad), rrows, nx);
    }
  }
  // obstacle flags of my boundary rows -> the neighbours' ghost rows (they mirror my rows)
  lbm::flags_ghost_push_kernel<<<64, 256, 0, s.stream>>>(
      s.flags, nx, s.rows, flags_of(s.lo_base, lo_ps) + (size_t)(G + lo_rows) * nx, flags_of(s.hi_base, hi_ps));
  CK(cudaGetLastError());
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
  if (world > MAX_WORLD) return fail("lbm_create_rank: at most %d ranks", MAX_WORLD);
  if (params->ny / world < G)
    return fail("lbm_create_rank: %d rows cannot be split into %d slabs of >= %d rows", params->ny, world, G);
  lbm_lattice* h = new_lattice(params, world);
  h->comm = new Comm();
  h->comm->rank = rank;
  h->comm->world = world;
  h->comm->halo = env_is("LBM_HALO", "nccl") ? HALO_NCCL : HALO_P2P;
  if (h->comm->halo == HALO_NCCL) h->reduce_per_step = false;
  if (h->stall_test) h->stall_test = env_int("LBM_TEST_RING_STALL", -1) == rank;
  set_ring_mode(h);
  h->slabs.resize(1);
  Slab& s = h->slabs[0];
  s.device = device;
  s.rank = rank;
  lbm_slab_rows(params->ny, world, rank, &s.y0, &s.rows);
  h->host_y0 = s.y0;
  if (create_slab(h, s, obstacles_slab, s.y0, s.rows, &h->tot_cells) || attach_comm(h, rank, world, nccl_unique_id)) {
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
    for (void* p : c->peers)
      if (p) cudaIpcCloseMemHandle(p);
    if (c->dummy_ghost) cudaFree(c->dummy_ghost);
    if (c->scratch64) cudaFree(c->scratch64);
    if (c->nccl && n) n->CommDestroy(c->nccl);
    delete c;
  }
  for (auto& s : h->slabs) {
    cudaSetDevice(s.device);
    drop_graphs(s);
    if (s.base) cudaFree(s.base);
    if (s.partials) cudaFree(s.partials);
    if (s.totals) cudaFree(s.totals);
    if (s.counter) cudaFree(s.counter);
    if (s.macro) cudaFree(s.macro);
    if (s.trace) cudaFree(s.trace);
    if (s.inbox) cudaFree(s.inbox);
    if (s.ev_begin) cudaEventDestroy(s.ev_begin);
    if (s.ev_end) cudaEventDestroy(s.ev_end);
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
    const long long cells = (long long)(s.rows + 2 * G) * h->p.nx;
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
      CK(cudaMemcpyAsync(s.buf[h->cur] + k * s.ps + (size_t)G * nx, cells[k] + (size_t)(s.y0 - h->host_y0) * nx,
                         sizeof(float) * (size_t)s.rows * nx, cudaMemcpyHostToDevice, s.stream));
  }
  if (sync_all(h)) return 1;      // the ghost zones are refreshed at the start of the next run
  return comm_barrier(h);
}

int lbm_download(lbm_lattice* h, float* const cells[9])
{
  if (!h || !cells) return fail("null argument");
  const int nx = h->p.nx;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    for (int k = 0; k < 9; k++)
      CK(cudaMemcpyAsync(cells[k] + (size_t)(s.y0 - h->host_y0) * nx, s.buf[h->cur] + k * s.ps + (size_t)G * nx,
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

namespace {
// sum of per-block partials of a whole-slab reduction kernel, over slabs and (rank mode) ranks
int finish_scalar(lbm_lattice* h, double local, double* out)
{
  double total = local;
  if (h->comm) {
    Slab& s = h->slabs[0];
    double* d = reinterpret_cast<double*>(h->comm->scratch64 + 1);
    CK(cudaMemcpyAsync(d, &total, sizeof(double), cudaMemcpyHostToDevice, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    NK(nccl_api()->AllReduce(d, d, 1, ncclDouble, ncclSum, h->comm->nccl, s.stream));
    CK(cudaMemcpyAsync(&total, d, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
  }
  *out = total;
  return 0;
}
}  // namespace

int lbm_av_velocity(lbm_lattice* h, float* av_vel)
{
  if (!h || !av_vel) return fail("null argument");
  double total = 0;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    const int nblk = std::min(s.np, 148 * 8);
    lbm::av_velocity_kernel<256><<<nblk, 256, 0, s.stream>>>(s.buf[h->cur] + (size_t)G * h->p.nx,
                                                           s.flags + (size_t)G * h->p.nx, s.ps, h->p.nx,
                                                           s.rows, s.partials);
    CK(cudaGetLastError());
    std::vector<double> part((size_t)nblk);
    CK(cudaMemcpyAsync(part.data(), s.partials, sizeof(double) * nblk, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (double v : part) total += v;
  }
  if (finish_scalar(h, total, &total)) return 1;
  *av_vel = (float)(total / (double)h->tot_cells);
  return 0;
}

int lbm_total_density(lbm_lattice* h, double* total_out)
{
  if (!h || !total_out) return fail("null argument");
  double total = 0;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    const int nblk = std::min(s.np, 148 * 8);
    lbm::total_density_kernel<256><<<nblk, 256, 0, s.stream>>>(s.buf[h->cur] + (size_t)G * h->p.nx, s.ps,
                                                             h->p.nx, s.rows, s.partials);
    CK(cudaGetLastError());
    std::vector<double> part((size_t)nblk);
    CK(cudaMemcpyAsync(part.data(), s.partials, sizeof(double) * nblk, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (double v : part) total += v;
  }
  return finish_scalar(h, total, total_out);
}

int lbm_macroscopic(lbm_lattice* h, float* ux, float* uy, float* speed, float* pressure)
{
  if (!h || !ux || !uy || !speed || !pressure) return fail("null argument");
  const int nx = h->p.nx;
  for (auto& s : h->slabs) {
    CK(cudaSetDevice(s.device));
    const size_t n = (size_t)s.rows * nx;
    if (!s.macro) CK(cudaMalloc(&s.macro, sizeof(float) * 4 * n));     // kept for the next call
    lbm::macroscopic_kernel<<<148 * 8, 256, 0, s.stream>>>(s.buf[h->cur] + (size_t)G * nx, s.flags + (size_t)G * nx,
                                                         s.ps, nx, s.rows, h->p.density, s.macro, s.macro + n,
                                                         s.macro + 2 * n, s.macro + 3 * n);
    CK(cudaGetLastError());
    float* outs[4] = {ux, uy, speed, pressure};
    for (int k = 0; k < 4; k++)
      CK(cudaMemcpyAsync(outs[k] + (size_t)(s.y0 - h->host_y0) * nx, s.macro + k * n, sizeof(float) * n,
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

int lbm_probe_l2_copy(unsigned long long bytes, int reps, double* gbs)
{
  if (!gbs || bytes < 4096 || reps < 1) return fail("lbm_probe_l2_copy: bad arguments");
  float4 *a = nullptr, *b = nullptr;
  const long long n4 = (long long)(bytes / 16);
  CK(cudaMalloc(&a, (size_t)n4 * 16));
  CK(cudaMalloc(&b, (size_t)n4 * 16));
  CK(cudaMemset(a, 0, (size_t)n4 * 16));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  lbm::l2_copy_probe_kernel<<<148 * 8, 256>>>(a, b, n4, 2);          // warm-up: both buffers into L2
  float best = 0;
  for (int k = 0; k < 3; k++) {
    CK(cudaEventRecord(e0));
    lbm::l2_copy_probe_kernel<<<148 * 8, 256>>>(a, b, n4, reps);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = (k == 0 || ms < best) ? ms : best;
  }
  *gbs = 2.0 * (double)n4 * 16.0 * reps / (best / 1e3) / 1e9;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
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
