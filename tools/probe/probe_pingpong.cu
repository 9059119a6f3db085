// probe_pingpong.cu -- how long does one word take from SM to SM?  (tuning aid for lbm_resident.cuh)
//   1. through L2: st.relaxed.gpu / ld.relaxed.gpu of a 64-bit word, two blocks on different SMs
//   2. same with st.release.gpu / ld.acquire.gpu
//   3. through distributed shared memory: 2-CTA cluster, st.shared::cluster into the partner, poll own smem
//   4. dependent ld.relaxed.gpu chain (L2 load latency as this access flavour sees it)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_pingpong probe_pingpong.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long ldr(const unsigned long long* p)
{ unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void str(unsigned long long* p, unsigned long long v)
{ asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lda(const unsigned long long* p)
{ unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void strel(unsigned long long* p, unsigned long long v)
{ asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// flavours of "publish one 64-bit word" / "look at it"
//  0: st.relaxed.gpu / ld.relaxed.gpu      1: st.release.gpu / ld.acquire.gpu   2: atom.exch / ld.relaxed.gpu
//  3: st.relaxed.gpu / atom.or(0)          4: atom.exch / atom.or(0)            5: st.volatile / ld.volatile
//  6: st.relaxed.sys / ld.relaxed.sys      7: st.relaxed.gpu / ld.global.cv     8: red.max / ld.relaxed.gpu
template <int F> __device__ __forceinline__ void pub(unsigned long long* p, unsigned long long v)
{
  if (F == 0 || F == 3 || F == 7) str(p, v);
  else if (F == 1) strel(p, v);
  else if (F == 2 || F == 4) atomicExch(p, v);
  else if (F == 5) *reinterpret_cast<volatile unsigned long long*>(p) = v;
  else if (F == 6) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  else if (F == 8) asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <int F> __device__ __forceinline__ unsigned long long look(unsigned long long* p)
{
  if (F == 0 || F == 2 || F == 8) return ldr(p);
  if (F == 1) return lda(p);
  if (F == 3 || F == 4) return atomicOr(p, 0ull);
  if (F == 5) return *reinterpret_cast<volatile unsigned long long*>(p);
  if (F == 6) { unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
  unsigned long long v; asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}

// blocks 0 and `peer` play; everybody else exits.  words: a at w[0], b at w[32] (different lines)
template <int F>
__global__ void pingpong_l2(unsigned long long* w, int peer, int n, long long* cycles, unsigned* smid)
{
  if (threadIdx.x != 0) return;
  const int me = blockIdx.x == 0 ? 0 : (blockIdx.x == peer ? 1 : -1);
  if (me < 0) return;
  unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); smid[me] = sm;
  unsigned long long* mine = w + (me ? 32 : 0);
  unsigned long long* theirs = w + (me ? 0 : 32);
  const long long t0 = clock64();
  for (int i = 1; i <= n; i++) {
    if (me == 0) {
      pub<F>(mine, i);
      while (look<F>(theirs) != (unsigned long long)i) {}
    } else {
      while (look<F>(theirs) != (unsigned long long)i) {}
      pub<F>(mine, i);
    }
  }
  if (me == 0) *cycles = clock64() - t0;
}

// one-way latency seen by a poller that keeps `depth` polls in flight (staggered): block 0 publishes i and
// waits for the echo; block `peer` polls with `depth` lanes of one warp, each lane delayed by lane * gap cycles
__global__ void pingpong_multi(unsigned long long* w, int peer, int n, int depth, long long* cycles)
{
  const int me = blockIdx.x == 0 ? 0 : (blockIdx.x == peer ? 1 : -1);
  if (me < 0 || threadIdx.x >= depth) return;
  unsigned long long* mine = w + (me ? 32 : 0);
  unsigned long long* theirs = w + (me ? 0 : 32);
  const unsigned mask = depth >= 32 ? 0xffffffffu : ((1u << depth) - 1u);
  const long long t0 = clock64();
  // de-phase the lanes once: lane l starts l * (300 / depth) cycles late, then every lane polls back to back
  const long long wait_until = t0 + (long long)threadIdx.x * (300 / depth);
  while (clock64() < wait_until) {}
  for (int i = 1; i <= n; i++) {
    if (me == 0 && threadIdx.x == 0) str(mine, i);
    // any lane seeing the value releases the warp
    for (;;) {
      const bool seen = ldr(theirs) >= (unsigned long long)i;
      if (__any_sync(mask, seen)) break;
    }
    if (me == 1 && threadIdx.x == 0) str(mine, i);
  }
  if (me == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

__global__ void __cluster_dims__(2, 1, 1) pingpong_dsmem(int n, long long* cycles)
{
  __shared__ unsigned long long box;
  cg::cluster_group cl = cg::this_cluster();
  const unsigned me = cl.block_rank();
  if (threadIdx.x == 0) box = 0;
  cl.sync();
  if (threadIdx.x == 0) {
    unsigned long long* theirs = cl.map_shared_rank(&box, me ^ 1);
    volatile unsigned long long* mine = &box;
    const long long t0 = clock64();
    for (int i = 1; i <= n; i++) {
      if (me == 0) { *theirs = i; while (*mine != (unsigned long long)i) {} }
      else         { while (*mine != (unsigned long long)i) {} *theirs = i; }
    }
    if (me == 0) *cycles = clock64() - t0;
  }
  cl.sync();
}

__global__ void chase(unsigned long long* w, int n, long long* cycles)
{
  unsigned long long i = 0;
  const long long t0 = clock64();
  for (int k = 0; k < n; k++) i = ldr(w + i);
  *cycles = clock64() - t0 + (long long)(i & 1);
}

int main()
{
  unsigned long long* w; long long* cyc; unsigned* smid;
  cudaMalloc(&w, 1 << 20); cudaMalloc(&cyc, 8); cudaMallocManaged(&smid, 8);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int n = 20000;
  long long c = 0;
  const char* names[] = {"st.relaxed.gpu / ld.relaxed.gpu", "st.release.gpu / ld.acquire.gpu", "atom.exch / ld.relaxed.gpu",
                         "st.relaxed.gpu / atom.or 0", "atom.exch / atom.or 0", "st.volatile / ld.volatile",
                         "st.relaxed.sys / ld.relaxed.sys", "st.relaxed.gpu / ld.cv", "red.max / ld.relaxed.gpu"};
  for (int peer : {1, 37, 74, 147}) {
    for (int f = 0; f < 9; f++) {
      cudaMemset(w, 0, 1 << 20);
      switch (f) {
        case 0: pingpong_l2<0><<<148, 32>>>(w, peer, n, cyc, smid); break;
        case 1: pingpong_l2<1><<<148, 32>>>(w, peer, n, cyc, smid); break;
        case 2: pingpong_l2<2><<<148, 32>>>(w, peer, n, cyc, smid); break;
        case 3: pingpong_l2<3><<<148, 32>>>(w, peer, n, cyc, smid); break;
        case 4: pingpong_l2<4><<<148, 32>>>(w, peer, n, cyc, smid); break;
        case 5: pingpong_l2<5><<<148, 32>>>(w, peer, n, cyc, smid); break;
        case 6: pingpong_l2<6><<<148, 32>>>(w, peer, n, cyc, smid); break;
        case 7: pingpong_l2<7><<<148, 32>>>(w, peer, n, cyc, smid); break;
        default: pingpong_l2<8><<<148, 32>>>(w, peer, n, cyc, smid); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("L2 ping-pong block 0 (SM %u) <-> block %d (SM %u)  %-34s %5.0f cycles per round trip (%s)\n", smid[0], peer,
             smid[1], names[f], (double)c / n, cudaGetErrorString(e));
    }
    for (int depth : {1, 2, 4, 8}) {
      cudaMemset(w, 0, 1 << 20);
      pingpong_multi<<<148, 32>>>(w, peer, n, depth, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("L2 ping-pong block 0 <-> block %d, %d staggered pollers per side: %5.0f cycles per round trip (%s)\n", peer, depth,
             (double)c / n, cudaGetErrorString(e));
    }
  }
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    pingpong_dsmem<<<2, 32>>>(n, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DSMEM ping-pong (2-CTA cluster): %.0f cycles, %.3f us per round trip (%s)\n", (double)c / n, ms * 1e3 / n,
           cudaGetErrorString(e));
  }
  {
    // a chain over 4096 words, 128 bytes apart
    static unsigned long long host[1 << 17];
    for (int i = 0; i < 4096; i++) host[i * 16] = (unsigned long long)(((i + 1) % 4096) * 16);
    cudaMemcpy(w, host, sizeof host, cudaMemcpyHostToDevice);
    chase<<<1, 1>>>(w, 4096, cyc);           // warm L2
    chase<<<1, 1>>>(w, n, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent ld.relaxed.gpu chain: %.0f cycles per load (%s); SM clock attribute %d kHz\n", (double)c / n,
           cudaGetErrorString(e), clk);
  }
  return 0;
}
