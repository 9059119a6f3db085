// lbm_resident.cuh -- the whole `for tt` loop (d2q9-bgk.c:203-234) of a SMALL lattice as ONE persistent
// launch with the state resident in shared memory (sm_100a).  Per timestep it does what lbm_step_kernel
// does -- accelerate_flow kernels.cl:7-42 (folded into the previous step's store), propagate :80-98,
// rebound :100-107, collision :109-196, av_velocity :198 + d2q9-bgk.c:408-423 -- with the same
// f32-strict arithmetic (bgk_cell), so the state is bit-identical to the one-step kernel and the oracle.
//
// Why: the shipped 128x128, 128x256 and 256x256 cases are 16 K - 64 K cells.  One timestep of them is
// ~0.3 us of work for 148 SMs, but a launch chained to the previous one costs ~2.5 us even from a CUDA
// graph with programmatic dependent launch (profiles/r1_tuning.md), so the one-step kernel ran those
// cases at 7-25 % of the HBM roofline.  Measured on the way here (profiles/r2_resident.md): a persistent
// kernel that keeps the state in L2 and orders neighbouring blocks with st.release / ld.acquire flags
// is no faster than launches (2.2-2.5 us per step: MEMBAR.ALL.GPU + flag + CCTL.IVALL round trips).
//
// So: the grid is launched once (cooperatively: every block is resident), block b owns R whole rows
// and keeps all nine populations of them in shared memory for the whole launch (two copies, ping-pong,
// ONE bar.sync per timestep), and the only thing that crosses SMs per step is the halo: the three
// populations the row above / below pulls (kernels.cl:92-98) of the block's last / first row.  Each
// halo value travels as ONE 64-bit word {step tag : float bits} written with st.relaxed.gpu into the
// neighbour's inbox in L2 and polled there with ld.relaxed.gpu -- data and flag in the same single-copy
// atomic word, so there is no fence, no separate flag and no second round trip (the scheme NCCL's LL
// protocol uses between GPUs, here between SMs).  Inboxes are double-buffered by step parity: a block
// can only be one step ahead of its neighbour, because it needs the neighbour's halo of step s to
// compute step s+1.  Boundary rows are relaxed and sent first, interior rows next, the poll last.
//
// Critical path of a step: relax the boundary rows -> 64-bit store -> L2 (~850 cycles until the neighbour's
// poll sees it, whatever the access flavour: tools/probe/probe_pingpong.cu) -> bar.sync.  Measured on B200:
// 128x128 2.78 -> 1.26 us per step, 128x256 2.78 -> 1.26, 256x256 2.84 -> 1.60 (profiles/r2_resident.md).
// Spins are bounded (2^22 polls, ~1 s) and raise *timed_out instead of hanging the GPU.
#pragma once
#include "lbm_kernels.cuh"

namespace lbm {

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// wait for the halo words of one or two cells (three per cell, one per pulled population) carrying `tag`.
// All loads of a poll are in flight together (one L2 round trip per poll, ~300 cycles); the poll count
// bounds the wait at ~1 s, and a time-out raised by anybody ends everybody's waits.
template <bool TWO>
__device__ __forceinline__ void halo_wait(const unsigned long long* pa, const unsigned long long* pb, int plane_stride,
                                          unsigned tag, unsigned* timed_out, float (&a)[3], float (&b)[3])
{
  unsigned long long va0, va1, va2, vb0 = 0, vb1 = 0, vb2 = 0;
  unsigned n = 0;
  for (;;) {
    va0 = ld_relaxed_gpu_u64(pa);
    va1 = ld_relaxed_gpu_u64(pa + plane_stride);
    va2 = ld_relaxed_gpu_u64(pa + 2 * plane_stride);
    unsigned all = ((unsigned)(va0 >> 32) ^ tag) | ((unsigned)(va1 >> 32) ^ tag) | ((unsigned)(va2 >> 32) ^ tag);
    if (TWO) {
      vb0 = ld_relaxed_gpu_u64(pb);
      vb1 = ld_relaxed_gpu_u64(pb + plane_stride);
      vb2 = ld_relaxed_gpu_u64(pb + 2 * plane_stride);
      all |= ((unsigned)(vb0 >> 32) ^ tag) | ((unsigned)(vb1 >> 32) ^ tag) | ((unsigned)(vb2 >> 32) ^ tag);
    }
    if (all == 0u) break;
    if ((++n & 1023u) == 0u) {
      if (*reinterpret_cast<volatile unsigned*>(timed_out) != 0u) break;
      if (n >= (1u << 22)) { *timed_out = 1u; break; }
    }
  }
  a[0] = __uint_as_float((unsigned)va0); a[1] = __uint_as_float((unsigned)va1); a[2] = __uint_as_float((unsigned)va2);
  b[0] = __uint_as_float((unsigned)vb0); b[1] = __uint_as_float((unsigned)vb1); b[2] = __uint_as_float((unsigned)vb2);
}
__device__ __forceinline__ void halo_send(unsigned long long* p, float v, unsigned tag)
{
  st_relaxed_gpu_u64(p, ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v));
}

struct ResidentArgs {
  const float*   src;        // state at launch start: plane k at src + k*ps, lattice row y at + (GHOST+y)*nx
  float*         dst;        // receives the state after nsteps (same geometry, the other buffer)
  const uint8_t* flags;      // same row indexing
  long long      ps;
  int            nx, ny;
  int            R;          // rows per block: block b owns rows [b*R, min(ny, b*R + R))
  int            nsteps;
  int            fuse_after; // another timestep of the same run follows this launch
  int            np;         // partial-sum slots per step: partials[t * np + block]
  float          omega, a1, a2;
  double*        partials;
  unsigned long long* inbox; // [gridDim.x][2 from-below/from-above][2 parities][3 planes][nx]
  unsigned       base;       // halo tags of this launch are base + 1 .. base + nsteps - 1
  unsigned*      timed_out;
  int            stall_block;// negative test of the time-out: this block never sends (-1 = none)
#ifdef LBM_RES_TRACE
  long long*     trace;      // tuning build only: clock64 at 5 points of steps 64..79, threads 0 and last of block 5
#endif
};

#ifdef LBM_RES_TRACE
#define RES_MARK(k)                                                                                   \
  if (A.trace && b == 5 && (tid == 0 || tid == nthr - 1) && t >= 64 && t < 80)                        \
    A.trace[((tid ? 1 : 0) * 16 + (t - 64)) * 6 + (k)] = clock64();
#else
#define RES_MARK(k)
#endif

constexpr int RES_BATCH = 8;     // timesteps whose speed sums are reduced together
constexpr int RES_MAX_CPT = 8;   // cells per thread the kernel is instantiated for: 1, 2, 4, 8

// shared memory of a block (dynamic): two state copies of (9R + 6) rows -- planes 2,5,6 carry the row
// below the block, planes 4,7,8 the row above -- and RES_BATCH per-thread speed sums (the cells' flags
// live in registers)
__host__ __device__ constexpr size_t resident_smem_bytes(int R, int nx, int threads)
{
  return 2 * sizeof(float) * (size_t)(9 * R + 6) * nx + sizeof(double) * RES_BATCH * (size_t)threads;
}

// CPT = cells per thread (thread `tid` owns cells tid, tid + nthr, ... of the block's boundary-rows-first
// order).  Everything about a thread's cells that does not change from step to step -- indices, the
// periodic x neighbours, which halo it feeds -- is worked out once, before the time loop: with one warp
// per scheduler every instruction of a step is on the critical path.
template <int CPT>
__global__ void __launch_bounds__(512, 1)
lbm_resident_kernel(const __grid_constant__ ResidentArgs A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nx = A.nx, R = A.R;
  const int b = blockIdx.x, nblk = gridDim.x;
  const int y0 = b * R;
  const int nr = min(R, A.ny - y0);
  const int ncell = nr * nx;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

  // plane k of a copy starts at off[k]; lattice row y of the block (-1 .. nr) lives at row y + o[k]
  // (o = 1 for the planes that carry the row below, folded into off[k]; 0 otherwise)
  const int buf_floats = (9 * R + 6) * nx;
  float* const buf0 = reinterpret_cast<float*>(smem_raw);
  double* const spd = reinterpret_cast<double*>(smem_raw + 2 * sizeof(float) * (size_t)buf_floats);
  int off[9];
  {
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const bool ghost = (k == 2 || k == 5 || k == 6 || k == 4 || k == 7 || k == 8);
      off[k] = acc + ((k == 2 || k == 5 || k == 6) ? nx : 0);
      acc += (R + (ghost ? 1 : 0)) * nx;
    }
  }

  // inbox[blk][dir][parity][j][x]: dir 0 = from below (planes 2,5,6), 1 = from above (planes 4,7,8)
  const int below = (b + nblk - 1) % nblk, above = (b + 1) % nblk;
  const size_t box = (size_t)12 * nx;
  unsigned long long* const send_up = A.inbox + (size_t)above * box;              // + par*3nx + j*nx + x
  unsigned long long* const send_dn = A.inbox + (size_t)below * box + 6 * nx;
  const unsigned long long* const recv = A.inbox + (size_t)b * box;               // + dir*6nx + par*3nx + j*nx + x

  // ---- this thread's cells.  Cells are numbered boundary rows first (virtual row 0 -> row 0, 1 -> row
  // nr-1, v -> row v-1), so that the halo is on its way while the interior is relaxed.  Kept as BYTE
  // offsets of the cell and of its periodic west / east neighbours inside a plane, so that a shared-memory
  // access of the time loop is [uniform plane base + one of these registers].
  int c4[CPT], w4[CPT], e4[CPT];
  unsigned cfl[CPT];            // the cell's flags; bit 8: first row of the block, bit 9: last row
#pragma unroll
  for (int j = 0; j < CPT; j++) {
    const int i = tid + j * nthr;
    c4[j] = -1; w4[j] = e4[j] = 0; cfl[j] = 0;
    if (i < ncell) {
      const int vr = i / nx, x = i - vr * nx;
      const int y = vr == 0 ? 0 : (vr == 1 ? nr - 1 : vr - 1);
      const int c = y * nx + x;
      c4[j] = 4 * c;
      w4[j] = 4 * (c + (x == 0 ? nx - 1 : -1));
      e4[j] = 4 * (c + (x == nx - 1 ? 1 - nx : 1));
      cfl[j] = (unsigned)A.flags[(long long)(GHOST + y0) * nx + c] | (y == 0 ? 256u : 0u) | (y == nr - 1 ? 512u : 0u);
    }
  }
  const int last_row4 = 4 * (nr - 1) * nx;
  auto at = [](const float* base, int byte_off) -> const float& {
    return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + byte_off);
  };
  auto to = [](float* base, int byte_off) -> float& {
    return *reinterpret_cast<float*>(reinterpret_cast<char*>(base) + byte_off);
  };

  // ---- load the block's rows and the two rows around it
  {
    const long long row0 = (long long)(GHOST + y0) * nx;
    for (int i = tid; i < ncell; i += nthr) {
#pragma unroll
      for (int k = 0; k < 9; k++) buf0[off[k] + i] = __ldcg(A.src + k * A.ps + row0 + i);
    }
    const long long row_lo = (long long)(GHOST + (y0 + A.ny - 1) % A.ny) * nx;
    const long long row_hi = (long long)(GHOST + (y0 + nr) % A.ny) * nx;
    for (int x = tid; x < nx; x += nthr) {
      buf0[off[2] - nx + x] = __ldcg(A.src + 2 * A.ps + row_lo + x);
      buf0[off[5] - nx + x] = __ldcg(A.src + 5 * A.ps + row_lo + x);
      buf0[off[6] - nx + x] = __ldcg(A.src + 6 * A.ps + row_lo + x);
      buf0[off[4] + nr * nx + x] = __ldcg(A.src + 4 * A.ps + row_hi + x);
      buf0[off[7] + nr * nx + x] = __ldcg(A.src + 7 * A.ps + row_hi + x);
      buf0[off[8] + nr * nx + x] = __ldcg(A.src + 8 * A.ps + row_hi + x);
    }
  }
  __syncthreads();

  for (int t = 0; t < A.nsteps; t++) {
    const float* cur = buf0 + (t & 1) * buf_floats;
    float* nxt = buf0 + ((t + 1) & 1) * buf_floats;
    const bool last = t + 1 == A.nsteps;
    const bool fuse = !last || A.fuse_after != 0;
    const unsigned tag = A.base + (unsigned)t + 1u;
    const int par3 = ((t + 1) & 1) * 3 * nx;
    const bool send = !last && b != A.stall_block;

    // the halo slots of this step (parity folded in)
    unsigned long long* const up_t = send_up + par3;
    unsigned long long* const dn_t = send_dn + par3;

    RES_MARK(0)
    double speed_sum = 0.0;
#pragma unroll
    for (int j = 0; j < CPT; j++) {
      const int c = c4[j];
      if (c < 0) continue;
      const int w = w4[j], e = e4[j];
      float f[9];
      f[0] = at(cur + off[0], c);
      f[1] = at(cur + off[1], w);
      f[2] = at(cur + off[2] - nx, c);
      f[3] = at(cur + off[3], e);
      f[4] = at(cur + off[4] + nx, c);
      f[5] = at(cur + off[5] - nx, w);
      f[6] = at(cur + off[6] - nx, e);
      f[7] = at(cur + off[7] + nx, e);
      f[8] = at(cur + off[8] + nx, w);
      const unsigned fl = cfl[j];
      if (fl & FLAG_OBSTACLE) {
        swap2(f[1], f[3]); swap2(f[2], f[4]); swap2(f[5], f[7]); swap2(f[6], f[8]);
      } else {
        const float usq = bgk_cell(f, A.omega);
        if (usq > 0.0f) speed_sum += (double)__fsqrt_rn(usq);
        if (fuse && (fl & FLAG_ACCEL)) accelerate_cell(f, A.a1, A.a2);
      }
      if (!last) {
        if (send) {
          if (fl & 512u) {     // last row: the block above pulls 2,5,6 from it
            unsigned long long* p = up_t + (unsigned)(c - last_row4) / 4u;
            halo_send(p, f[2], tag); halo_send(p + nx, f[5], tag); halo_send(p + 2 * nx, f[6], tag);
          }
          if (fl & 256u) {     // first row: the block below pulls 4,7,8 from it
            unsigned long long* p = dn_t + (unsigned)c / 4u;
            halo_send(p, f[4], tag); halo_send(p + nx, f[7], tag); halo_send(p + 2 * nx, f[8], tag);
          }
        }
#pragma unroll
        for (int k = 0; k < 9; k++) to(nxt + off[k], c) = f[k];
      } else {
        float* o = A.dst + (long long)(GHOST + y0) * nx + c / 4;
#pragma unroll
        for (int k = 0; k < 9; k++) __stcg(o + k * A.ps, f[k]);
      }
    }
    RES_MARK(1)
    // per-thread speed sum of the step; the block sum is taken for RES_BATCH steps at a time (below)
    spd[(t & (RES_BATCH - 1)) * nthr + tid] = speed_sum;

    // ---- the two rows around the block for the next step, from the inboxes: item i < nx is column i of
    // the row below, item nx + i column i of the row above; a thread waits for two items at a time
    if (!last) {
      for (int i0 = tid; i0 < 2 * nx; i0 += 2 * nthr) {
        const int i1 = i0 + nthr;
        const bool two = i1 < 2 * nx;
        const int d0 = i0 >= nx, d1 = i1 >= nx;
        const int x0 = i0 - d0 * nx, x1 = i1 - d1 * nx;
        const unsigned long long* pa = recv + d0 * 6 * nx + par3 + x0;
        float va[3], vb[3];
        if (two) halo_wait<true>(pa, recv + d1 * 6 * nx + par3 + x1, nx, tag, A.timed_out, va, vb);
        else halo_wait<false>(pa, pa, nx, tag, A.timed_out, va, vb);
        if (d0 == 0) { nxt[off[2] - nx + x0] = va[0]; nxt[off[5] - nx + x0] = va[1]; nxt[off[6] - nx + x0] = va[2]; }
        else { nxt[off[4] + nr * nx + x0] = va[0]; nxt[off[7] + nr * nx + x0] = va[1]; nxt[off[8] + nr * nx + x0] = va[2]; }
        if (two) {
          if (d1 == 0) { nxt[off[2] - nx + x1] = vb[0]; nxt[off[5] - nx + x1] = vb[1]; nxt[off[6] - nx + x1] = vb[2]; }
          else { nxt[off[4] + nr * nx + x1] = vb[0]; nxt[off[7] + nr * nx + x1] = vb[1]; nxt[off[8] + nr * nx + x1] = vb[2]; }
        }
      }
    }
    RES_MARK(2)
    __syncthreads();
    RES_MARK(3)

    // ---- block sums of the last RES_BATCH steps' speeds, off the per-step critical path: warp w takes
    // steps w, w + nwarp, ...; lanes add the threads' values in a fixed order, then a fixed shuffle tree
    if ((t & (RES_BATCH - 1)) == RES_BATCH - 1 || last) {
      const int nb = (t & (RES_BATCH - 1)) + 1, t_first = t - (nb - 1);
      for (int sidx = warp; sidx < nb; sidx += nwarp) {
        double v = 0.0;
        for (int j = lane; j < nthr; j += 32) v += spd[sidx * nthr + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL_MASK, v, o);
        if (lane == 0) A.partials[(long long)(t_first + sidx) * A.np + b] = v;
      }
      __syncthreads();     // the next batch overwrites spd
    }
    RES_MARK(4)
  }
}

}  // namespace lbm
