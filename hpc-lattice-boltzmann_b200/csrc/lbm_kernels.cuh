// lbm_kernels.cuh -- sm_100a device code for the D2Q9-BGK per-timestep path.
//
// Replaces, in ONE pass per timestep, what the reference does in two OpenCL kernels plus a
// host round trip (reference file:line):
//   accelerate_flow  kernels.cl:7-42      -> epilogue of the previous step (or accelerate_row_kernel)
//   propagate        kernels.cl:80-98     -> periodic pull, 128-bit loads + warp shuffles
//   rebound          kernels.cl:100-107   -> register permutation, selected per cell
//   collision        kernels.cl:109-196   -> f32-strict BGK relaxation (arithmetic contract below)
//   av_velocity      kernels.cl:198 + d2q9-bgk.c:408-423 -> fused double-precision tree reduction
//
// Layout (DESIGN.md "Data layout"): SoA, 9 planes per buffer, two buffers (ping-pong).  A slab
// of `rows` lattice rows is stored with GHOST ghost rows below and above; the kernels of this file
// work on a view whose row 0 is the nearest ghost row below and whose row rows+1 is the nearest one
// above, so the y-periodic wrap and the multi-GPU halo are the same mechanism: whoever computes a
// boundary row also stores the three populations its vertical neighbour will pull into that
// neighbour's ghost row (which is this GPU's own ghost row when there is one GPU, or a peer pointer
// over NVLink when there are several).  The streaming kernel (lbm_stream.cuh) uses all GHOST rows.
//
// Arithmetic contract ("f32-strict"; mirrored bit-for-bit by oracle/canon_impl.h VARIANT_B200):
// every operation below is written as an explicit round-to-nearest intrinsic, so nvcc can neither
// contract nor reassociate it.  State after any number of steps is bit-identical to the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lbm {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int FLAG_OBSTACLE = 1;   // cell is blocked (d2q9-bgk.c:627)
constexpr int FLAG_ACCEL    = 2;   // fluid cell of global row ny-2 (kernels.cl:21,29)
constexpr int GHOST = 4;           // ghost rows below and above a slab's owned rows (all nine planes)

// system-scope flag accesses for the cross-GPU ring ordering
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin until *flag >= want (wrap-safe); gives up after ~4 s of SM clocks and raises
// *timed_out instead of hanging the GPU -- the host turns that into an error
// (once the flag is raised later waits give up at once: a dead neighbour costs one time-out, not one per launch)
__device__ __forceinline__ void spin_until(const unsigned* flag, unsigned want, unsigned* timed_out)
{
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - want) < 0) {
    if (*reinterpret_cast<volatile unsigned*>(timed_out) != 0u) break;
    if (clock64() - t0 > 8000000000LL) { *timed_out = 1u; break; }
    __nanosleep(100);
  }
}

// LBM_REDUCE=step -- the north-star's per-step allreduce of the speed sum (it replaces the serial
// host sum d2q9-bgk.c:416-423), done by the step kernel itself: the last block of a launch to
// finish adds up the launch's per-block partials in a fixed order and stores the slab total of
// each timestep into EVERY rank's table (peer stores over NVLink); a rank then adds the table's
// rows in rank order.  No collective launch, no host involvement, deterministic.
constexpr int MAX_REDUCE_WORLD = 16;
struct StepReduce {
  double*   peer[MAX_REDUCE_WORLD];   // every rank's table [world][cap] (own included), null = off
  int       world, rank, cap;
  int       slot;                     // table column of this launch's first timestep
  unsigned* ticket;                   // blocks of this launch that are done
};

// called by every thread of every block after the block's partials are written (by its threads
// 0 .. nsteps-1).  partials: [nsteps][stride], `count` valid entries per step.  scratch: >= NT doubles
// of shared memory.
template <int NT>
__device__ __forceinline__ void last_block_allreduce(const StepReduce& R, const double* partials, int stride,
                                                     int count, int nsteps, double* scratch)
{
  if (R.peer[0] == nullptr) return;
  __shared__ int is_last;
  // only the threads that wrote partials fence (a fence by every thread would make each block wait
  // for all of its own population stores to drain before it may leave the SM)
  if (threadIdx.x < nsteps) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(R.ticket, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int s = 0; s < nsteps; s++) {
    double v = 0.0;
    for (int i = threadIdx.x; i < count; i += NT) v += __ldcg(partials + (long long)s * stride + i);
    scratch[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      double t = 0.0;
      for (int i = threadIdx.x; i < NT; i += 32) t += scratch[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(FULL_MASK, t, off);
      if (threadIdx.x == 0) {
        const int col = (R.slot + s) % R.cap;
        for (int r = 0; r < R.world; r++) R.peer[r][(long long)R.rank * R.cap + col] = t;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { *R.ticket = 0u; __threadfence_system(); }
}

struct StepArgs {
  const float*   src;        // 9 planes, plane k at src + k*ps, storage row r at + r*nx
  float*         dst;
  const uint8_t* flags;      // per storage cell, same row indexing as a plane
  long long      ps;         // plane stride (floats)
  long long      nvec;       // rows * nxv work items
  int            nx, rows, nxv;
  unsigned       div_mul;    // item / nxv == umulhi(item, div_mul) >> div_shift for item < 2^31
  int            div_shift;  // (nxv == 1: div_mul == 0 means the quotient is the item itself)
  float          omega, a1, a2;
  int            fuse_accel; // apply the NEXT step's accelerate_flow to the values being stored
  float*         ghost_lo[3];// row base receiving planes 4,7,8 of the first owned row
  float*         ghost_hi[3];// row base receiving planes 2,5,6 of the last owned row
  double*        partials;   // [gridDim.x] per-block sums of cell speeds for this step
  // one-process-per-GPU ring ordering done by the boundary blocks themselves (null = not used):
  const unsigned* ring_in;   // [0] steps finished by my lower neighbour's top row, [1] by my upper's bottom row
  unsigned*      ring_out_lo;// lower neighbour's ring_in[1]
  unsigned*      ring_out_hi;// upper neighbour's ring_in[0]
  unsigned*      ring_tickets;// [2] boundary blocks of this step that are done (lo side, hi side)
  unsigned*      ring_timeout;
  unsigned       ring_step;  // number of steps every rank has completed before this one
  int            rot;        // block-id rotation: the blocks holding the last row run first
  int            nb_lo, nb_hi;// how many blocks touch the first / the last owned row
  StepReduce     red;        // LBM_REDUCE=step (peer[0] == null: off)
};

// ---- vector access helpers ---------------------------------------------------------------
template <int VEC> __device__ __forceinline__ void ld_vec(const float* p, float (&v)[VEC]);
template <> __device__ __forceinline__ void ld_vec<1>(const float* p, float (&v)[1]) { v[0] = *p; }
template <> __device__ __forceinline__ void ld_vec<2>(const float* p, float (&v)[2]) {
  const float2 a = *reinterpret_cast<const float2*>(p); v[0] = a.x; v[1] = a.y;
}
template <> __device__ __forceinline__ void ld_vec<4>(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <int VEC> __device__ __forceinline__ void st_vec(float* p, const float (&v)[VEC]);
template <> __device__ __forceinline__ void st_vec<1>(float* p, const float (&v)[1]) { *p = v[0]; }
template <> __device__ __forceinline__ void st_vec<2>(float* p, const float (&v)[2]) {
  *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}
template <> __device__ __forceinline__ void st_vec<4>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

template <int VEC> __device__ __forceinline__ unsigned ld_flags(const uint8_t* p);
template <> __device__ __forceinline__ unsigned ld_flags<1>(const uint8_t* p) { return *p; }
template <> __device__ __forceinline__ unsigned ld_flags<2>(const uint8_t* p) {
  return *reinterpret_cast<const uint16_t*>(p);
}
template <> __device__ __forceinline__ unsigned ld_flags<4>(const uint8_t* p) {
  return *reinterpret_cast<const uint32_t*>(p);
}

// ---- one fluid cell: BGK relaxation in place (f32-strict contract) ----------------------------
// t[] holds the pulled populations on entry and the relaxed ones on return; returns |u|^2 of the
// pre-collision moments (kernels.cl:198 takes the speed from the same moments).
__device__ __forceinline__ float bgk_cell(float (&t)[9], float omega)
{
  constexpr float W0 = (float)(4.0 / 9.0);    // float roundings of the double quotients,
  constexpr float W1 = (float)(1.0 / 9.0);    // as kernels.cl:58-61 has them
  constexpr float W2 = (float)(1.0 / 36.0);

  float rho = __fadd_rn(t[0], t[1]);
  rho = __fadd_rn(rho, t[2]); rho = __fadd_rn(rho, t[3]); rho = __fadd_rn(rho, t[4]);
  rho = __fadd_rn(rho, t[5]); rho = __fadd_rn(rho, t[6]); rho = __fadd_rn(rho, t[7]);
  rho = __fadd_rn(rho, t[8]);
  const float mx = __fsub_rn(__fadd_rn(__fadd_rn(t[1], t[5]), t[8]),
                             __fadd_rn(__fadd_rn(t[3], t[6]), t[7]));
  const float my = __fsub_rn(__fadd_rn(__fadd_rn(t[2], t[5]), t[6]),
                             __fadd_rn(__fadd_rn(t[4], t[7]), t[8]));
  // one correctly rounded reciprocal and two products instead of the reference's two divisions:
  // rho is O(0.1), so the reciprocal never leaves the fast path, whereas mx/rho with mx == 0
  // (fluid at rest: most of a freshly started channel) takes the slow IEEE-division path
  const float inv = __frcp_rn(rho);
  const float ux = __fmul_rn(mx, inv);
  const float uy = __fmul_rn(my, inv);
  const float usq = __fmaf_rn(uy, uy, __fmul_rn(ux, ux));
  const float b = __fmaf_rn(-1.5f, usq, 1.0f);
  const float wr0 = __fmul_rn(W0, rho), wr1 = __fmul_rn(W1, rho), wr2 = __fmul_rn(W2, rho);
  const float u5 = __fadd_rn(ux, uy), u6 = __fsub_rn(uy, ux);

  t[0] = __fmaf_rn(omega, __fsub_rn(__fmul_rn(wr0, b), t[0]), t[0]);
#define LBM_RELAX(k, u, wr)                                                          \
  {                                                                                  \
    const float p = __fmaf_rn((u), __fmaf_rn((u), 4.5f, 3.0f), b);                   \
    t[k] = __fmaf_rn(omega, __fsub_rn(__fmul_rn((wr), p), t[k]), t[k]);              \
  }
  LBM_RELAX(1,  ux, wr1) LBM_RELAX(2,  uy, wr1) LBM_RELAX(3, -ux, wr1) LBM_RELAX(4, -uy, wr1)
  LBM_RELAX(5,  u5, wr2) LBM_RELAX(6,  u6, wr2) LBM_RELAX(7, -u5, wr2) LBM_RELAX(8, -u6, wr2)
#undef LBM_RELAX
  return usq;
}

// inflow acceleration (kernels.cl:29-41) of one fluid cell of row ny-2, on registers
__device__ __forceinline__ void accelerate_cell(float (&t)[9], float a1, float a2)
{
  if (__fsub_rn(t[3], a1) > 0.0f && __fsub_rn(t[6], a2) > 0.0f && __fsub_rn(t[7], a2) > 0.0f) {
    t[1] = __fadd_rn(t[1], a1); t[5] = __fadd_rn(t[5], a2); t[8] = __fadd_rn(t[8], a2);
    t[3] = __fsub_rn(t[3], a1); t[6] = __fsub_rn(t[6], a2); t[7] = __fsub_rn(t[7], a2);
  }
}

__device__ __forceinline__ void swap2(float& a, float& b) { const float c = a; a = b; b = c; }

// block-wide deterministic sum (fixed shuffle tree, then fixed order over warps)
template <int TPB>
__device__ __forceinline__ double block_sum(double v)
{
  __shared__ double warp_part[TPB / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(FULL_MASK, v, off);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (warp == 0) {
    s = lane < TPB / 32 ? warp_part[lane] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(FULL_MASK, s, off);
  }
  return s;   // valid in thread 0
}

// ---- the fused timestep --------------------------------------------------------------------
// One thread = VEC consecutive cells of one row.  Grid = ceil(rows*nx/VEC / TPB) blocks.
template <int VEC, int TPB>
__global__ void __launch_bounds__(TPB, 1024 / TPB)   // 64 registers: 1024 resident threads per SM
lbm_step_kernel(const __grid_constant__ StepArgs A)
{
  // Programmatic dependent launch: let the NEXT step's grid be scheduled as soon as every block of
  // this one is resident; its blocks park at griddepcontrol.wait (below) until this grid has
  // completed and flushed, so launch latency and block ramp-up hide under the previous step's tail.
  asm volatile("griddepcontrol.launch_dependents;");
  // the blocks that hold the slab's last row are rotated to the front of the grid (rot = how many),
  // so both boundary rows are computed -- and their halos are on the wire -- first
  const unsigned vb = A.rot == 0 ? blockIdx.x
                      : (blockIdx.x < (unsigned)A.rot ? gridDim.x - A.rot + blockIdx.x : blockIdx.x - A.rot);
  bool ring_lo = false, ring_hi = false;
  if (A.ring_in != nullptr) {
    // Only the blocks that touch a boundary row take part in the cross-GPU ordering.  Before
    // step s they need "neighbour's adjacent boundary row has finished step s-1": its stores into
    // my ghost row (my input) have landed and it no longer reads the ghost row I will overwrite.
    const unsigned first = vb * TPB;
    const unsigned last = min(first + TPB, (unsigned)A.nvec) - 1u;
    ring_lo = first < (unsigned)A.nxv;
    ring_hi = last >= (unsigned)(A.rows - 1) * (unsigned)A.nxv;
    if (ring_lo || ring_hi) {
      if (threadIdx.x == 0) {
        if (ring_lo) spin_until(A.ring_in + 0, A.ring_step, A.ring_timeout);
        if (ring_hi) spin_until(A.ring_in + 1, A.ring_step, A.ring_timeout);
      }
      __syncthreads();
    }
  }
  const unsigned gid = vb * TPB + threadIdx.x;             // nvec < 2^31 is checked at create
  const bool active = gid < (unsigned)A.nvec;
  const unsigned item = active ? gid : (unsigned)A.nvec - 1u;   // idle tail threads shadow the last item
  const unsigned q = A.div_mul ? (__umulhi(item, A.div_mul) >> A.div_shift) : item;
  const int r = (int)q + 1;                                // storage row (1..rows)
  const int c = (int)(item - q * (unsigned)A.nxv);
  const int x = c * VEC;
  const int lane = threadIdx.x & 31;
  // the element just outside my vector comes from the neighbouring lane when that lane holds the
  // adjacent cells of the same row; otherwise (warp edge, row edge, periodic wrap) I load it
  const bool west_by_load = !(lane > 0 && c > 0);
  const bool east_by_load = !(lane < 31 && c < A.nxv - 1);
  const int xw = (x == 0) ? A.nx - 1 : x - 1;
  const int xe = (x + VEC == A.nx) ? 0 : x + VEC;

  const long long row_mid = (long long)r * A.nx;
  const long long row_lo = row_mid - A.nx, row_hi = row_mid + A.nx;
  const long long ps = A.ps;
  const float* s = A.src;
  const float* p1 = s + 1 * ps + row_mid; const float* p3 = s + 3 * ps + row_mid;
  const float* p5 = s + 5 * ps + row_lo;  const float* p6 = s + 6 * ps + row_lo;
  const float* p7 = s + 7 * ps + row_hi;  const float* p8 = s + 8 * ps + row_hi;

  // ---- phase 1: issue EVERY global load of this thread before the first use, so that a warp
  // pays one memory round trip per step, not one per dependent group
  float f[9][VEC];
  float v1[VEC], v3[VEC], v5[VEC], v6[VEC], v7[VEC], v8[VEC];
  float e1 = 0.f, e3 = 0.f, e5 = 0.f, e6 = 0.f, e7 = 0.f, e8 = 0.f;
  const unsigned flags = ld_flags<VEC>(A.flags + row_mid + x);   // constant data: may precede the wait
  // everything below reads what the previous step wrote (and overwrites what it read)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  ld_vec<VEC>(s + row_mid + x, f[0]);
  ld_vec<VEC>(p1 + x, v1);
  ld_vec<VEC>(s + 2 * ps + row_lo + x, f[2]);
  ld_vec<VEC>(p3 + x, v3);
  ld_vec<VEC>(s + 4 * ps + row_hi + x, f[4]);
  ld_vec<VEC>(p5 + x, v5);
  ld_vec<VEC>(p6 + x, v6);
  ld_vec<VEC>(p7 + x, v7);
  ld_vec<VEC>(p8 + x, v8);
  if (west_by_load) { e1 = p1[xw]; e5 = p5[xw]; e8 = p8[xw]; }
  if (east_by_load) { e3 = p3[xe]; e6 = p6[xe]; e7 = p7[xe]; }

  // ---- phase 2: assemble the x-shifted populations (speeds 1,5,8 come from x-1; 3,6,7 from x+1)
  {
    const float l1 = __shfl_up_sync(FULL_MASK, v1[VEC - 1], 1);
    const float l5 = __shfl_up_sync(FULL_MASK, v5[VEC - 1], 1);
    const float l8 = __shfl_up_sync(FULL_MASK, v8[VEC - 1], 1);
    const float r3 = __shfl_down_sync(FULL_MASK, v3[0], 1);
    const float r6 = __shfl_down_sync(FULL_MASK, v6[0], 1);
    const float r7 = __shfl_down_sync(FULL_MASK, v7[0], 1);
    f[1][0] = west_by_load ? e1 : l1;
    f[5][0] = west_by_load ? e5 : l5;
    f[8][0] = west_by_load ? e8 : l8;
    f[3][VEC - 1] = east_by_load ? e3 : r3;
    f[6][VEC - 1] = east_by_load ? e6 : r6;
    f[7][VEC - 1] = east_by_load ? e7 : r7;
#pragma unroll
    for (int i = 1; i < VEC; i++) { f[1][i] = v1[i - 1]; f[5][i] = v5[i - 1]; f[8][i] = v8[i - 1]; }
#pragma unroll
    for (int i = 0; i < VEC - 1; i++) { f[3][i] = v3[i + 1]; f[6][i] = v6[i + 1]; f[7][i] = v7[i + 1]; }
  }

  // ---- phase 3: per cell, rebound (obstacle) or BGK relaxation (+ next step's acceleration)
  double speed_sum = 0.0;
  {
    // (a thread-level "all my cells are plain fluid" fast path was measured 5-25 % SLOWER on B200,
    // profiles/r1_tuning.md, so there is one generic loop)
    const bool fuse = A.fuse_accel != 0;
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      const unsigned fl = (flags >> (8 * j)) & 0xffu;
      float t[9];
#pragma unroll
      for (int k = 0; k < 9; k++) t[k] = f[k][j];
      if (fl & FLAG_OBSTACLE) {
        // rebound (kernels.cl:100-107): opposite directions of the pulled values, rest population
        // kept.  A real branch instead of the reference's 0/1 multiply (:179-196): a zero-density
        // obstacle cell cannot leak a NaN.
        swap2(t[1], t[3]); swap2(t[2], t[4]); swap2(t[5], t[7]); swap2(t[6], t[8]);
      } else {
        const float usq = bgk_cell(t, A.omega);
        if (usq > 0.0f) speed_sum += (double)__fsqrt_rn(usq);
        // inflow acceleration of the NEXT step on the just-relaxed values of row ny-2:
        // bit-identical to running accelerate_flow as a separate pre-pass (kernels.cl:7-42)
        if (fuse && (fl & FLAG_ACCEL)) accelerate_cell(t, A.a1, A.a2);
      }
#pragma unroll
      for (int k = 0; k < 9; k++) f[k][j] = t[k];
    }
  }

  if (active) {
    float* d = A.dst + row_mid + x;
#pragma unroll
    for (int k = 0; k < 9; k++) st_vec<VEC>(d + k * ps, f[k]);
    if (r == 1) {            // my south neighbour pulls 4,7,8 from this row
      st_vec<VEC>(A.ghost_lo[0] + x, f[4]);
      st_vec<VEC>(A.ghost_lo[1] + x, f[7]);
      st_vec<VEC>(A.ghost_lo[2] + x, f[8]);
    }
    if (r == A.rows) {       // my north neighbour pulls 2,5,6 from this row
      st_vec<VEC>(A.ghost_hi[0] + x, f[2]);
      st_vec<VEC>(A.ghost_hi[1] + x, f[5]);
      st_vec<VEC>(A.ghost_hi[2] + x, f[6]);
    }
  } else {
    speed_sum = 0.0;
  }

  if (ring_lo || ring_hi) {
    // my stores (local rows and the neighbour's ghost row over NVLink) are ordered before the
    // ticket; the last boundary block of a side publishes "step ring_step done" to that neighbour
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (ring_lo && atomicAdd(A.ring_tickets + 0, 1u) == (unsigned)A.nb_lo - 1u) {
        A.ring_tickets[0] = 0u;
        __threadfence_system();
        st_release_sys(A.ring_out_lo, A.ring_step + 1u);
      }
      if (ring_hi && atomicAdd(A.ring_tickets + 1, 1u) == (unsigned)A.nb_hi - 1u) {
        A.ring_tickets[1] = 0u;
        __threadfence_system();
        st_release_sys(A.ring_out_hi, A.ring_step + 1u);
      }
    }
  }

  const double total = block_sum<TPB>(speed_sum);
  if (threadIdx.x == 0) A.partials[vb] = total;
  if (A.red.peer[0] != nullptr) {
    __shared__ double red_scratch[TPB];
    last_block_allreduce<TPB>(A.red, A.partials, 0, (int)gridDim.x, 1, red_scratch);
  }
}

// ---- helpers shared with the streaming kernel (lbm_stream.cuh) ----------------------------------
// relax the 4 cells a thread holds; speeds are accumulated only when `count`
template <bool FAST>
__device__ __forceinline__ double relax_vec4(float (&f)[9][4], unsigned flags, float omega, float a1,
                                             float a2, bool fuse, bool count)
{
  double sum = 0.0;
  if (FAST && (flags & 0x03030303u) == 0u) {
    // plain fluid everywhere: four independent relaxations the scheduler can interleave
    float usq[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float t[9];
#pragma unroll
      for (int k = 0; k < 9; k++) t[k] = f[k][j];
      usq[j] = bgk_cell(t, omega);
#pragma unroll
      for (int k = 0; k < 9; k++) f[k][j] = t[k];
    }
    if (count) {
#pragma unroll
      for (int j = 0; j < 4; j++) if (usq[j] > 0.0f) sum += (double)__fsqrt_rn(usq[j]);
    }
    return sum;
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned fl = (flags >> (8 * j)) & 0xffu;
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; k++) t[k] = f[k][j];
    if (fl & FLAG_OBSTACLE) {
      swap2(t[1], t[3]); swap2(t[2], t[4]); swap2(t[5], t[7]); swap2(t[6], t[8]);
    } else {
      const float usq = bgk_cell(t, omega);
      if (count && usq > 0.0f) sum += (double)__fsqrt_rn(usq);
      if (fuse && (fl & FLAG_ACCEL)) accelerate_cell(t, a1, a2);
    }
#pragma unroll
    for (int k = 0; k < 9; k++) f[k][j] = t[k];
  }
  return sum;
}

__device__ __forceinline__ float4 f4(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }

// ---- small kernels ---------------------------------------------------------------------------

// stand-alone accelerate_flow (kernels.cl:7-42) on storage row `r` of a buffer; used for the first
// step of a run (every later step gets it from the previous step's epilogue).  The ghost zones are
// refreshed right after it, so it touches the owned row only.
__global__ void accelerate_row_kernel(float* buf, const uint8_t* flags, long long ps, int nx, int r,
                                      float a1, float a2)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nx) return;
  const long long i = (long long)r * nx + x;
  if (!(flags[i] & FLAG_ACCEL)) return;
  float* f1 = buf + 1 * ps + i; float* f3 = buf + 3 * ps + i; float* f5 = buf + 5 * ps + i;
  float* f6 = buf + 6 * ps + i; float* f7 = buf + 7 * ps + i; float* f8 = buf + 8 * ps + i;
  if (__fsub_rn(*f3, a1) > 0.0f && __fsub_rn(*f6, a2) > 0.0f && __fsub_rn(*f7, a2) > 0.0f) {
    *f1 = __fadd_rn(*f1, a1); *f5 = __fadd_rn(*f5, a2); *f8 = __fadd_rn(*f8, a2);
    *f3 = __fsub_rn(*f3, a1); *f6 = __fsub_rn(*f6, a2); *f7 = __fsub_rn(*f7, a2);
  }
}

// initial equilibrium fill on the device (d2q9-bgk.c:573-594), ghost rows included
__global__ void init_equilibrium_kernel(float* buf, long long ps, long long cells, float w0,
                                        float w1, float w2)
{
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
    buf[i] = w0;
#pragma unroll
    for (int k = 1; k <= 4; k++) buf[k * ps + i] = w1;
#pragma unroll
    for (int k = 5; k <= 8; k++) buf[k * ps + i] = w2;
  }
}

// second stage of the average-velocity reduction: block s sums the `count` partials of step s in a
// fixed order and writes the slab's speed total for that step (divided later by tot_cells).
// `counter` holds the index of the first step of this chunk inside `totals`; `stride` = partial
// slots per step.
__global__ void reduce_partials_kernel(const double* partials, int stride, int count, double* totals,
                                       const long long* counter)
{
  __shared__ double sm[256];
  const double* p = partials + (long long)blockIdx.x * stride;
  double v = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) v += p[i];
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[*counter + blockIdx.x] = sm[0];
}

__global__ void advance_counter_kernel(long long* counter, int by) { *counter += by; }

// av_velocity on a resident state (d2q9-bgk.c:426-475): per-block sums of cell speeds computed
// from the stored populations (no streaming, no collision).  Feeds calc_reynolds (:815-820).
template <int TPB>
__global__ void __launch_bounds__(TPB)
av_velocity_kernel(const float* buf, const uint8_t* flags, long long ps, int nx, int rows,
                   double* partials)
{
  const long long n = (long long)rows * nx;
  const long long stride = (long long)gridDim.x * TPB;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) {
    const long long q = i;                      // buf / flags point at the first owned row
    if (flags[q] & FLAG_OBSTACLE) continue;
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; k++) t[k] = buf[k * ps + q];
    float rho = __fadd_rn(t[0], t[1]);
#pragma unroll
    for (int k = 2; k < 9; k++) rho = __fadd_rn(rho, t[k]);
    const float mx = __fsub_rn(__fadd_rn(__fadd_rn(t[1], t[5]), t[8]),
                               __fadd_rn(__fadd_rn(t[3], t[6]), t[7]));
    const float my = __fsub_rn(__fadd_rn(__fadd_rn(t[2], t[5]), t[6]),
                               __fadd_rn(__fadd_rn(t[4], t[7]), t[8]));
    const float ux = __fdiv_rn(mx, rho), uy = __fdiv_rn(my, rho);
    acc += (double)__fsqrt_rn(__fmaf_rn(uy, uy, __fmul_rn(ux, ux)));
  }
  const double total = block_sum<TPB>(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

// total mass of a resident state (d2q9-bgk.c:822-838, the reference's DEBUG conservation check):
// per-block double sums over every population of every cell, obstacles included
template <int TPB>
__global__ void __launch_bounds__(TPB)
total_density_kernel(const float* buf, long long ps, int nx, int rows, double* partials)
{
  const long long n = (long long)rows * nx;
  const long long stride = (long long)gridDim.x * TPB;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) {
    const long long q = i;
#pragma unroll
    for (int k = 0; k < 9; k++) acc += (double)buf[k * ps + q];
  }
  const double total = block_sum<TPB>(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

// final-state fields of write_values (d2q9-bgk.c:857-897), same float expressions
__global__ void macroscopic_kernel(const float* buf, const uint8_t* flags, long long ps, int nx,
                                   int rows, float density, float* ux_out, float* uy_out,
                                   float* u_out, float* p_out)
{
  constexpr float C_SQ = (float)(1.0 / 3.0);
  const long long n = (long long)rows * nx;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long q = i;
    if (flags[q] & FLAG_OBSTACLE) {
      ux_out[i] = 0.0f; uy_out[i] = 0.0f; u_out[i] = 0.0f;
      p_out[i] = __fmul_rn(density, C_SQ);
      continue;
    }
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; k++) t[k] = buf[k * ps + q];
    float rho = __fadd_rn(0.0f, t[0]);
#pragma unroll
    for (int k = 1; k < 9; k++) rho = __fadd_rn(rho, t[k]);
    const float mx = __fsub_rn(__fadd_rn(__fadd_rn(t[1], t[5]), t[8]),
                               __fadd_rn(__fadd_rn(t[3], t[6]), t[7]));
    const float my = __fsub_rn(__fadd_rn(__fadd_rn(t[2], t[5]), t[6]),
                               __fadd_rn(__fadd_rn(t[4], t[7]), t[8]));
    const float ux = __fdiv_rn(mx, rho), uy = __fdiv_rn(my, rho);
    ux_out[i] = ux; uy_out[i] = uy;
    u_out[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)));
    p_out[i] = __fmul_rn(rho, C_SQ);
  }
}

// measurement aid (bench.py's L2 roofline for the 1024^2 case): `reps` passes of a read + write copy
// over two buffers small enough to live in L2, inside ONE launch (no launch overhead in the figure)
__global__ void l2_copy_probe_kernel(const float4* __restrict__ a, float4* __restrict__ b, long long n4, int reps)
{
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 v = __ldcg(a + i);
      v.x += (float)r;                 // keeps the compiler from collapsing the passes
      __stcg(b + i, v);
    }
  }
}

// ---- cross-GPU step ordering (one process per GPU: peers are other processes' memory mapped
// through CUDA IPC, so stream events cannot order them) ---------------------------------------
// Each rank owns two counters that its ring neighbours bump after every completed step.  Before
// step s a rank needs both counters >= s: the neighbours' stores into its ghost rows (the input of
// step s) have landed, and the neighbours no longer read the ghost rows it is about to overwrite.
// 2 threads: thread i waits for flags[i] >= want.  Bounded: after ~4 s of SM clocks it gives up
// and raises *timed_out instead of hanging the GPU (the host reports it as an error).
__global__ void wait_neighbours_kernel(const unsigned* flags, unsigned want, unsigned* timed_out)
{
  spin_until(flags + threadIdx.x, want, timed_out);
}

// runs after the step kernel in stream order: every store of that step (local and peer) is
// complete; publish "I have finished `done` steps" to both neighbours
__global__ void signal_neighbours_kernel(unsigned* peer_lo_flag, unsigned* peer_hi_flag, unsigned done)
{
  __threadfence_system();
  st_release_sys(threadIdx.x == 0 ? peer_lo_flag : peer_hi_flag, done);
}

}  // namespace lbm
