/*
 * oracle/canon.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the D2Q9-BGK per-timestep path of AlexDalt/HPC-Lattice-Boltzmann.
 * Three arithmetic variants are instantiated from canon_impl.h (see its header):
 *   f64_*      golden generator   (pinned byte-for-byte to the reference's check/ *.dat)
 *   f32ref_*   reference's float/double mixed arithmetic (kernels.cl)
 *   f32b200_*  the f32-strict operation order the CUDA kernel is specified to follow
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py compares the f64 variant's text
 * output with the sha256 of every golden file the reference ships (manifest in
 * tests/golden/), and regenerates the two golden files the checkout is missing.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -mfma -fopenmp).
 * As a library it exports <variant>_canon_{init,accelerate,step,run,av_velocity,macroscopic};
 * with -DCANON_MAIN it is a CLI:
 *   canon <f64|f32ref|f32b200> <paramfile> <obstaclefile> [outdir] [iters_override]
 * writing av_vels.dat / final_state.dat in the reference's formats (d2q9-bgk.c:900,915).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#define VARIANT_GOLD  1
#define VARIANT_REF32 2
#define VARIANT_B200  3
#define CANON_VARIANT_IDS

#define REAL double
#define SUFFIX(name) f64_##name
#define VARIANT VARIANT_GOLD
#include "canon_impl.h"
#undef REAL
#undef SUFFIX
#undef VARIANT

#define REAL float
#define SUFFIX(name) f32ref_##name
#define VARIANT VARIANT_REF32
#include "canon_impl.h"
#undef REAL
#undef SUFFIX
#undef VARIANT

#define REAL float
#define SUFFIX(name) f32b200_##name
#define VARIANT VARIANT_B200
#include "canon_impl.h"
#undef REAL
#undef SUFFIX
#undef VARIANT

#ifdef CANON_MAIN

static void fail(const char* msg)
{
  fprintf(stderr, "canon: %s\n", msg);
  exit(EXIT_FAILURE);
}

typedef struct { int nx, ny, iters, re_dim; double density, accel, omega; } case_t;

/* params file: 7 scalars (d2q9-bgk.c:499-525).  Read as double: the golden files were
 * produced by a double-precision code; float variants narrow afterwards. */
static case_t read_params(const char* path)
{
  case_t c;
  FILE* fp = fopen(path, "r");
  if (!fp) fail("cannot open param file");
  if (fscanf(fp, "%d %d %d %d %lf %lf %lf", &c.nx, &c.ny, &c.iters, &c.re_dim,
             &c.density, &c.accel, &c.omega) != 7) fail("bad param file");
  fclose(fp);
  return c;
}

/* obstacle file: "x y 1" per line, duplicates allowed (d2q9-bgk.c:615-628) */
static int* read_obstacles(const char* path, int nx, int ny)
{
  int* o = (int*)calloc((size_t)nx * ny, sizeof(int));
  FILE* fp = fopen(path, "r");
  if (!o || !fp) fail("cannot open obstacle file");
  int x, y, b, r;
  while ((r = fscanf(fp, "%d %d %d", &x, &y, &b)) != EOF) {
    if (r != 3 || x < 0 || x >= nx || y < 0 || y >= ny || b != 1) fail("bad obstacle line");
    o[(size_t)y * nx + x] = 1;
  }
  fclose(fp);
  return o;
}

static double now(void)
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + tv.tv_usec * 1e-6;
}

int main(int argc, char** argv)
{
  if (argc < 4) fail("usage: canon <f64|f32ref|f32b200> <paramfile> <obstaclefile> [outdir] [iters]");
  const char* variant = argv[1];
  case_t c = read_params(argv[2]);
  int* obst = read_obstacles(argv[3], c.nx, c.ny);
  const char* outdir = argc > 4 ? argv[4] : ".";
  if (argc > 5) c.iters = atoi(argv[5]);
  const size_t n = (size_t)c.nx * c.ny;
  double* av = (double*)malloc(sizeof(double) * (size_t)c.iters);
  char path[4096];
  FILE* fp;
  double t0, t1;

  snprintf(path, sizeof path, "%s/final_state.dat", outdir);
  fp = fopen(path, "w");
  if (!fp) fail("cannot open final_state.dat for writing");

  if (!strcmp(variant, "f64")) {
    double* f = (double*)malloc(sizeof(double) * 9 * n);
    double* m = (double*)malloc(sizeof(double) * 4 * n);
    f64_canon_init(c.nx, c.ny, c.density, f);
    t0 = now();
    f64_canon_run(c.nx, c.ny, c.density, c.accel, c.omega, obst, f, c.iters, av);
    t1 = now();
    f64_canon_macroscopic(c.nx, c.ny, c.density, obst, f, m, m + n, m + 2 * n, m + 3 * n);
    for (int y = 0; y < c.ny; y++)
      for (int x = 0; x < c.nx; x++) {
        const size_t k = (size_t)y * c.nx + x;
        fprintf(fp, "%d %d %.12E %.12E %.12E %.12E %d\n", x, y, m[k], m[n + k], m[2 * n + k],
                m[3 * n + k], obst[k]);
      }
    const double visc = 1.0 / 6.0 * (2.0 / c.omega - 1.0);
    printf("Reynolds number:\t\t%.12E\n",
           f64_canon_av_velocity(c.nx, c.ny, obst, f) * c.re_dim / visc);
  } else {
    const int strict = !strcmp(variant, "f32b200");
    if (!strict && strcmp(variant, "f32ref")) fail("unknown variant");
    float* f = (float*)malloc(sizeof(float) * 9 * n);
    float* m = (float*)malloc(sizeof(float) * 4 * n);
    const float d = (float)c.density, a = (float)c.accel, w = (float)c.omega;
    t0 = now();
    if (strict) {
      f32b200_canon_init(c.nx, c.ny, d, f);
      f32b200_canon_run(c.nx, c.ny, d, a, w, obst, f, c.iters, av);
      f32b200_canon_macroscopic(c.nx, c.ny, d, obst, f, m, m + n, m + 2 * n, m + 3 * n);
    } else {
      f32ref_canon_init(c.nx, c.ny, d, f);
      f32ref_canon_run(c.nx, c.ny, d, a, w, obst, f, c.iters, av);
      f32ref_canon_macroscopic(c.nx, c.ny, d, obst, f, m, m + n, m + 2 * n, m + 3 * n);
    }
    t1 = now();
    for (int y = 0; y < c.ny; y++)
      for (int x = 0; x < c.nx; x++) {
        const size_t k = (size_t)y * c.nx + x;
        fprintf(fp, "%d %d %.12E %.12E %.12E %.12E %d\n", x, y, m[k], m[n + k], m[2 * n + k],
                m[3 * n + k], obst[k]);
      }
    /* the float variants narrow av_vels like the reference's float av_vels[] does */
    for (int t = 0; t < c.iters; t++) av[t] = (double)(float)av[t];
  }
  fclose(fp);

  snprintf(path, sizeof path, "%s/av_vels.dat", outdir);
  fp = fopen(path, "w");
  if (!fp) fail("cannot open av_vels.dat for writing");
  for (int t = 0; t < c.iters; t++) fprintf(fp, "%d:\t%.12E\n", t, av[t]);
  fclose(fp);

  printf("Elapsed time:\t\t\t%.6lf (s)\n", t1 - t0);
  printf("MLUPS:\t\t\t\t%.3f\n", (double)n * c.iters / (t1 - t0) / 1e6);
  return 0;
}
#endif /* CANON_MAIN */
