/*
 * oracle/ref_shim/shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Host-memory implementation of the OpenCL entry points declared in CL/opencl.h.
 * The reference's device code (kernels.cl) is compiled as C right here, unmodified,
 * found through -I<reference dir>:  `kernel` -> static, `global` -> nothing,
 * get_global_id/get_global_size -> thread-private values set by the NDRange loop.
 * OpenCL C and C99 share the usual arithmetic conversions for everything that file
 * does, so (with -ffp-contract=off) this is the reference's own arithmetic.
 *
 * The NDRange loop is the only place with an OpenMP pragma: the reference sources
 * carry none, so "-fopenmp" parallelism of the CPU baseline lives in this dispatcher.
 * Work-items are independent (the kernels have no barriers, local memory or atomics),
 * hence the result is identical for any thread count.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "CL/opencl.h"

/* ---- the reference's kernels, compiled as plain C -------------------------------- */
static __thread size_t shim_gid[3];
static size_t shim_gsz[3];
static inline size_t get_global_id(unsigned d)   { return shim_gid[d]; }
static inline size_t get_global_size(unsigned d) { return shim_gsz[d]; }

#define kernel static
#define global
#include "kernels.cl"
#undef kernel
#undef global

/* ---- object model ------------------------------------------------------------------ */
enum { K_ACCELERATE = 1, K_COMP = 2, MAX_ARGS = 32 };

struct shim_platform { int unused; };
struct shim_device   { int unused; };
struct shim_context  { int unused; };
struct shim_queue    { int unused; };
struct shim_program  { int unused; };
struct shim_mem      { void* host; size_t bytes; };
struct shim_kernel {
  int which;
  union { void* p; int i; float f; char raw[8]; } arg[MAX_ARGS];
};

static struct shim_platform the_platform;
static struct shim_device   the_device;
static struct shim_context  the_context;
static struct shim_queue    the_queue;
static struct shim_program  the_program;

static void set_err(cl_int* e, cl_int v) { if (e) *e = v; }

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id* out, cl_uint* count)
{
  if (out && n) out[0] = &the_platform;
  if (count) *count = 1;
  return CL_SUCCESS;
}

cl_int clGetDeviceIDs(cl_platform_id p, cl_device_type t, cl_uint n, cl_device_id* out,
                      cl_uint* count)
{
  (void)p; (void)t;
  if (out && n) out[0] = &the_device;
  if (count) *count = 1;
  return CL_SUCCESS;
}

cl_int clGetDeviceInfo(cl_device_id d, cl_device_info what, size_t cap, void* dst, size_t* len)
{
  static const char name[] = "host-memory OpenCL shim (CPU, oracle only)";
  (void)d; (void)what;
  if (dst && cap) { strncpy((char*)dst, name, cap); ((char*)dst)[cap - 1] = 0; }
  if (len) *len = sizeof name;
  return CL_SUCCESS;
}

cl_context clCreateContext(const cl_context_properties* pr, cl_uint n, const cl_device_id* d,
                           void (*cb)(const char*, const void*, size_t, void*), void* ud,
                           cl_int* err)
{
  (void)pr; (void)n; (void)d; (void)cb; (void)ud;
  set_err(err, CL_SUCCESS);
  return &the_context;
}

cl_command_queue clCreateCommandQueue(cl_context c, cl_device_id d,
                                      cl_command_queue_properties p, cl_int* err)
{
  (void)c; (void)d; (void)p;
  set_err(err, CL_SUCCESS);
  return &the_queue;
}

cl_program clCreateProgramWithSource(cl_context c, cl_uint n, const char** src, const size_t* len,
                                     cl_int* err)
{
  (void)c; (void)n; (void)src; (void)len;   /* the source text is ignored: see top of file */
  set_err(err, CL_SUCCESS);
  return &the_program;
}

cl_int clBuildProgram(cl_program p, cl_uint n, const cl_device_id* d, const char* opts,
                      void (*cb)(cl_program, void*), void* ud)
{
  (void)p; (void)n; (void)d; (void)opts; (void)cb; (void)ud;
  return CL_SUCCESS;
}

cl_int clGetProgramBuildInfo(cl_program p, cl_device_id d, cl_program_build_info what, size_t cap,
                             void* dst, size_t* len)
{
  (void)p; (void)d; (void)what;
  if (dst && cap) ((char*)dst)[0] = 0;
  if (len) *len = 1;
  return CL_SUCCESS;
}

cl_kernel clCreateKernel(cl_program p, const char* name, cl_int* err)
{
  (void)p;
  struct shim_kernel* k = (struct shim_kernel*)calloc(1, sizeof *k);
  if (k && !strcmp(name, "accelerate_flow")) k->which = K_ACCELERATE;
  else if (k && !strcmp(name, "comp_func")) k->which = K_COMP;
  else { free(k); set_err(err, CL_INVALID_VALUE); return NULL; }
  set_err(err, CL_SUCCESS);
  return k;
}

cl_mem clCreateBuffer(cl_context c, cl_mem_flags fl, size_t bytes, void* host, cl_int* err)
{
  (void)c; (void)fl; (void)host;
  struct shim_mem* m = (struct shim_mem*)malloc(sizeof *m);
  if (m) { m->host = calloc(1, bytes ? bytes : 1); m->bytes = bytes; }
  if (!m || !m->host) { free(m); set_err(err, CL_INVALID_VALUE); return NULL; }
  set_err(err, CL_SUCCESS);
  return m;
}

cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t bytes, const void* val)
{
  if (!k || idx >= MAX_ARGS || bytes > sizeof k->arg[0].raw) return CL_INVALID_VALUE;
  memset(k->arg[idx].raw, 0, sizeof k->arg[idx].raw);
  memcpy(k->arg[idx].raw, val, bytes);
  return CL_SUCCESS;
}

/* a cl_mem argument was stored as the handle value; fetch its host storage */
#define BUF(T, k, i) ((T*)((struct shim_mem*)(k)->arg[i].p)->host)

static void run_accelerate(const struct shim_kernel* k)
{
  accelerate_flow(BUF(int, k, 0), k->arg[1].i, k->arg[2].i, k->arg[3].f, k->arg[4].f,
                  BUF(float, k, 5), BUF(float, k, 6), BUF(float, k, 7), BUF(float, k, 8),
                  BUF(float, k, 9), BUF(float, k, 10), BUF(float, k, 11), BUF(float, k, 12),
                  BUF(float, k, 13));
}

static void run_comp(const struct shim_kernel* k)
{
  comp_func(BUF(float, k, 0), BUF(int, k, 1), k->arg[2].i, k->arg[3].i, k->arg[4].f,
            BUF(float, k, 5), BUF(float, k, 6), BUF(float, k, 7), BUF(float, k, 8),
            BUF(float, k, 9), BUF(float, k, 10), BUF(float, k, 11), BUF(float, k, 12),
            BUF(float, k, 13), BUF(float, k, 14), BUF(float, k, 15), BUF(float, k, 16),
            BUF(float, k, 17), BUF(float, k, 18), BUF(float, k, 19), BUF(float, k, 20),
            BUF(float, k, 21), BUF(float, k, 22));
}

cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint dims, const size_t* off,
                              const size_t* gsz, const size_t* lsz, cl_uint nw,
                              const cl_event* wl, cl_event* ev)
{
  (void)q; (void)off; (void)lsz; (void)nw; (void)wl; (void)ev;
  if (!k || dims < 1 || dims > 2) return CL_INVALID_VALUE;
  const long g0 = (long)gsz[0];
  const long g1 = dims > 1 ? (long)gsz[1] : 1;
  shim_gsz[0] = (size_t)g0; shim_gsz[1] = (size_t)g1; shim_gsz[2] = 1;
  void (*body)(const struct shim_kernel*) = k->which == K_ACCELERATE ? run_accelerate : run_comp;
  if (dims == 1) {
#pragma omp parallel for schedule(static)
    for (long a = 0; a < g0; a++) {
      shim_gid[0] = (size_t)a; shim_gid[1] = 0; shim_gid[2] = 0;
      body(k);
    }
  } else {
#pragma omp parallel for schedule(static)
    for (long b = 0; b < g1; b++)
      for (long a = 0; a < g0; a++) {
        shim_gid[0] = (size_t)a; shim_gid[1] = (size_t)b; shim_gid[2] = 0;
        body(k);
      }
  }
  return CL_SUCCESS;
}

cl_int clFinish(cl_command_queue q) { (void)q; return CL_SUCCESS; }

cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem m, cl_bool blocking, size_t off,
                            size_t bytes, const void* src, cl_uint nw, const cl_event* wl,
                            cl_event* ev)
{
  (void)q; (void)blocking; (void)nw; (void)wl; (void)ev;
  if (!m || off + bytes > m->bytes) return CL_INVALID_VALUE;
  memcpy((char*)m->host + off, src, bytes);
  return CL_SUCCESS;
}

cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem m, cl_bool blocking, size_t off,
                           size_t bytes, void* dst, cl_uint nw, const cl_event* wl, cl_event* ev)
{
  (void)q; (void)blocking; (void)nw; (void)wl; (void)ev;
  if (!m || off + bytes > m->bytes) return CL_INVALID_VALUE;
  memcpy(dst, (const char*)m->host + off, bytes);
  return CL_SUCCESS;
}

cl_int clReleaseMemObject(cl_mem m) { if (m) { free(m->host); free(m); } return CL_SUCCESS; }
cl_int clReleaseKernel(cl_kernel k) { free(k); return CL_SUCCESS; }
cl_int clReleaseProgram(cl_program p) { (void)p; return CL_SUCCESS; }
cl_int clReleaseCommandQueue(cl_command_queue q) { (void)q; return CL_SUCCESS; }
cl_int clReleaseContext(cl_context c) { (void)c; return CL_SUCCESS; }
