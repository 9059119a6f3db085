/*
 * oracle/ref_shim/CL/opencl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Minimal host-memory stand-in for the OpenCL 1.x API, exposing exactly the types,
 * constants and 20 entry points that the reference's d2q9-bgk.c uses, so that the
 * reference translation unit compiles UNMODIFIED with gcc on a machine that has no
 * OpenCL headers, ICD or device.  "Device" buffers are host allocations; an NDRange
 * launch is a (optionally OpenMP-parallel) loop nest calling the reference's kernels,
 * which shim.c pulls in by #include "kernels.cl" compiled as C.
 */
#ifndef LBM_ORACLE_OPENCL_SHIM_H
#define LBM_ORACLE_OPENCL_SHIM_H

#include <stddef.h>
#include <stdint.h>

typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef float    cl_float;
typedef uint64_t cl_ulong;
typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef cl_uint  cl_device_info;
typedef cl_uint  cl_program_build_info;
typedef intptr_t cl_context_properties;

typedef struct shim_platform* cl_platform_id;
typedef struct shim_device*   cl_device_id;
typedef struct shim_context*  cl_context;
typedef struct shim_queue*    cl_command_queue;
typedef struct shim_program*  cl_program;
typedef struct shim_kernel*   cl_kernel;
typedef struct shim_mem*      cl_mem;
typedef struct shim_event*    cl_event;

#define CL_SUCCESS                 0
#define CL_BUILD_PROGRAM_FAILURE (-11)
#define CL_INVALID_VALUE         (-30)
#define CL_FALSE                   0
#define CL_TRUE                    1
#define CL_DEVICE_TYPE_ALL         0xFFFFFFFF
#define CL_DEVICE_NAME             0x102B
#define CL_MEM_READ_WRITE          (1 << 0)
#define CL_MEM_WRITE_ONLY          (1 << 1)
#define CL_PROGRAM_BUILD_LOG       0x1183

cl_int clGetPlatformIDs(cl_uint, cl_platform_id*, cl_uint*);
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint, cl_device_id*, cl_uint*);
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void*, size_t*);
cl_context clCreateContext(const cl_context_properties*, cl_uint, const cl_device_id*,
                           void (*)(const char*, const void*, size_t, void*), void*, cl_int*);
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties,
                                      cl_int*);
cl_program clCreateProgramWithSource(cl_context, cl_uint, const char**, const size_t*, cl_int*);
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id*, const char*,
                      void (*)(cl_program, void*), void*);
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t, void*,
                             size_t*);
cl_kernel clCreateKernel(cl_program, const char*, cl_int*);
cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void*, cl_int*);
cl_int clSetKernelArg(cl_kernel, cl_uint, size_t, const void*);
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel, cl_uint, const size_t*, const size_t*,
                              const size_t*, cl_uint, const cl_event*, cl_event*);
cl_int clFinish(cl_command_queue);
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, const void*,
                            cl_uint, const cl_event*, cl_event*);
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void*, cl_uint,
                           const cl_event*, cl_event*);
cl_int clReleaseMemObject(cl_mem);
cl_int clReleaseKernel(cl_kernel);
cl_int clReleaseProgram(cl_program);
cl_int clReleaseCommandQueue(cl_command_queue);
cl_int clReleaseContext(cl_context);

#endif
