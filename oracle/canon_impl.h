/*
 * oracle/canon_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's per-timestep D2Q9-BGK path.  This header is
 * included three times by canon.c, once per arithmetic variant:
 *
 *   VARIANT_GOLD    all-double canonical serial order (d2q9-bgk.c:78-82) with the
 *                   arithmetic of kernels.cl:17-41 (accelerate), :80-98 (propagate),
 *                   :100-107 (rebound), :109-196 (collision) and the post-collision
 *                   av_velocity of d2q9-bgk.c:435-474.  This is the generator of the
 *                   reference's golden files (check/ *.dat) -- pinned byte-for-byte.
 *   VARIANT_REF32   float state with the mixed float/double promotions that the
 *                   reference's OpenCL kernel actually performs (double literals in
 *                   kernels.cl:148-177, float constants :58-61, pre-collision speed
 *                   :198, sequential float host sum d2q9-bgk.c:416-423 with the
 *                   ii*ny+jj index quirk Q1 repaired).
 *   VARIANT_B200    the f32-strict operation order that the sm_100a kernel is
 *                   specified to execute (DESIGN.md "arithmetic contract"): every
 *                   multiply/add/fma below is one IEEE-754 binary32 operation, so
 *                   the GPU state must be BIT-IDENTICAL to this variant.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Required macros: REAL, SUFFIX(name), VARIANT (one of the three ids).
 */

#ifndef CANON_VARIANT_IDS
#define CANON_VARIANT_IDS
#define VARIANT_GOLD  1
#define VARIANT_REF32 2
#define VARIANT_B200  3
#endif

/* opposite-direction table: 1<->3, 2<->4, 5<->7, 6<->8 (kernels.cl:100-107) */
static const int SUFFIX(opp)[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};

/* initial equilibrium fill, d2q9-bgk.c:573-594 (every cell, obstacles included) */
void SUFFIX(canon_init)(int nx, int ny, REAL density, REAL* f)
{
  const size_t n = (size_t)nx * (size_t)ny;
  const REAL w0 = density * 4.0 / 9.0;
  const REAL w1 = density / 9.0;
  const REAL w2 = density / 36.0;
  for (size_t c = 0; c < n; c++) {
    f[0 * n + c] = w0;
    for (int k = 1; k <= 4; k++) f[(size_t)k * n + c] = w1;
    for (int k = 5; k <= 8; k++) f[(size_t)k * n + c] = w2;
  }
}

/* in-place inflow acceleration on row ny-2, kernels.cl:17-41 */
void SUFFIX(canon_accelerate)(int nx, int ny, REAL density, REAL accel,
                              const int* obst, REAL* f)
{
  const size_t n = (size_t)nx * (size_t)ny;
  const REAL a1 = density * accel / 9.0;   /* REAL*REAL, then a double divide, rounded to REAL */
  const REAL a2 = density * accel / 36.0;
  const size_t row = (size_t)(ny - 2) * (size_t)nx;
  REAL* f1 = f + 1 * n; REAL* f3 = f + 3 * n; REAL* f5 = f + 5 * n;
  REAL* f6 = f + 6 * n; REAL* f7 = f + 7 * n; REAL* f8 = f + 8 * n;
  for (int x = 0; x < nx; x++) {
    const size_t c = row + (size_t)x;
    if (!obst[c] && (f3[c] - a1) > 0.0 && (f6[c] - a2) > 0.0 && (f7[c] - a2) > 0.0) {
      f1[c] += a1; f5[c] += a2; f8[c] += a2;
      f3[c] -= a1; f6[c] -= a2; f7[c] -= a2;
    }
  }
}

/* one cell of the collision, per variant.  t[] = pulled populations, o[] = outputs,
 * returns the speed this cell contributes to the average */
static inline REAL SUFFIX(collide_cell)(const REAL* t, REAL* o, REAL omega)
{
#if VARIANT == VARIANT_GOLD
  const double c_sq = 1.0 / 3.0;
  const double W[9] = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0,
                       1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
  double ld = 0.0;
  for (int k = 0; k < 9; k++) ld += t[k];
  const double ux = (t[1] + t[5] + t[8] - (t[3] + t[6] + t[7])) / ld;
  const double uy = (t[2] + t[5] + t[6] - (t[4] + t[7] + t[8])) / ld;
  const double usq = ux * ux + uy * uy;
  const double u[9] = {0.0, ux, uy, -ux, -uy, ux + uy, -ux + uy, -ux - uy, ux - uy};
  for (int k = 0; k < 9; k++) {
    const double de = W[k] * ld * (1.0 + u[k] / c_sq + (u[k] * u[k]) / (2.0 * c_sq * c_sq)
                                   - usq / (2.0 * c_sq));
    o[k] = t[k] + omega * (de - t[k]);
  }
  /* canonical av_velocity looks at the post-collision state, d2q9-bgk.c:435-474 */
  double ld2 = 0.0;
  for (int k = 0; k < 9; k++) ld2 += o[k];
  const double vx = (o[1] + o[5] + o[8] - (o[3] + o[6] + o[7])) / ld2;
  const double vy = (o[2] + o[5] + o[6] - (o[4] + o[7] + o[8])) / ld2;
  return sqrt(vx * vx + vy * vy);

#elif VARIANT == VARIANT_REF32
  /* float constants are float roundings of the double quotients, kernels.cl:58-61 */
  const float c_sq = 1.0 / 3.0;
  const float w0 = 4.0 / 9.0, w1 = 1.0 / 9.0, w2 = 1.0 / 36.0;
  float ld = 0.0;
  for (int k = 0; k < 9; k++) ld += t[k];
  const float ux = (t[1] + t[5] + t[8] - (t[3] + t[6] + t[7])) / ld;
  const float uy = (t[2] + t[5] + t[6] - (t[4] + t[7] + t[8])) / ld;
  const float usq = ux * ux + uy * uy;
  const float u[9] = {0.0f, ux, uy, -ux, -uy, ux + uy, -ux + uy, -ux - uy, ux - uy};
  float de[9];
  /* the double literals promote the bracket and the final product; the store rounds to float */
  de[0] = w0 * ld * (1.0 - usq / (2.0 * c_sq));
  for (int k = 1; k < 9; k++) {
    const float w = (k < 5) ? w1 : w2;
    de[k] = w * ld * (1.0 + u[k] / c_sq + (u[k] * u[k]) / (2.0 * c_sq * c_sq)
                      - usq / (2.0 * c_sq));
  }
  for (int k = 0; k < 9; k++) o[k] = t[k] + omega * (de[k] - t[k]);
  /* kernels.cl:198 -- speed from the pre-collision moments, stored as float */
  return (float)sqrt((ux * ux) + (uy * uy));

#else /* VARIANT_B200: one IEEE binary32 operation per line item, fmaf == GPU FMA */
  const float W0 = (float)(4.0 / 9.0), W1 = (float)(1.0 / 9.0), W2 = (float)(1.0 / 36.0);
  float rho = t[0];
  for (int k = 1; k < 9; k++) rho = rho + t[k];
  const float mx = ((t[1] + t[5]) + t[8]) - ((t[3] + t[6]) + t[7]);
  const float my = ((t[2] + t[5]) + t[6]) - ((t[4] + t[7]) + t[8]);
  const float inv = 1.0f / rho;          /* correctly rounded reciprocal == __frcp_rn */
  const float ux = mx * inv;
  const float uy = my * inv;
  const float usq = fmaf(uy, uy, ux * ux);
  const float b = fmaf(-1.5f, usq, 1.0f);
  const float wr0 = W0 * rho, wr1 = W1 * rho, wr2 = W2 * rho;
  const float u5 = ux + uy, u6 = uy - ux;
  const float u[9] = {0.0f, ux, uy, -ux, -uy, u5, u6, -u5, -u6};
  o[0] = fmaf(omega, wr0 * b - t[0], t[0]);
  for (int k = 1; k < 9; k++) {
    const float p = fmaf(u[k], fmaf(u[k], 4.5f, 3.0f), b);
    const float e = ((k < 5) ? wr1 : wr2) * p;
    o[k] = fmaf(omega, e - t[k], t[k]);
  }
  return sqrtf(usq);
#endif
}

/*
 * One full timestep src -> dst: accelerate (in place on src), periodic pull, bounce-back
 * or BGK collision, average speed.  `speeds` is scratch of nx*ny REALs.
 * Returns av. velocity of this step as double (callers narrow it as their variant says).
 * tot_cells = number of non-obstacle cells (d2q9-bgk.c:146-152).
 */
double SUFFIX(canon_step)(int nx, int ny, REAL density, REAL accel, REAL omega,
                          const int* obst, REAL* src, REAL* dst, REAL* speeds,
                          long tot_cells, int do_accel)
{
  const size_t n = (size_t)nx * (size_t)ny;
  if (do_accel) SUFFIX(canon_accelerate)(nx, ny, density, accel, obst, src);

#pragma omp parallel for schedule(static)
  for (int y = 0; y < ny; y++) {
    const int yn = (y + 1) % ny;
    const int ys = (y == 0) ? ny - 1 : y - 1;
    for (int x = 0; x < nx; x++) {
      const int xe = (x + 1) % nx;
      const int xw = (x == 0) ? nx - 1 : x - 1;
      const size_t c = (size_t)y * nx + x;
      REAL t[9], o[9];
      /* pull, kernels.cl:90-98 */
      t[0] = src[0 * n + (size_t)y  * nx + x ];
      t[1] = src[1 * n + (size_t)y  * nx + xw];
      t[2] = src[2 * n + (size_t)ys * nx + x ];
      t[3] = src[3 * n + (size_t)y  * nx + xe];
      t[4] = src[4 * n + (size_t)yn * nx + x ];
      t[5] = src[5 * n + (size_t)ys * nx + xw];
      t[6] = src[6 * n + (size_t)ys * nx + xe];
      t[7] = src[7 * n + (size_t)yn * nx + xe];
      t[8] = src[8 * n + (size_t)yn * nx + xw];
      if (obst[c]) {
        /* rebound: opposite directions of the pulled values; rest population kept */
        for (int k = 0; k < 9; k++) o[k] = t[SUFFIX(opp)[k]];
        speeds[c] = 0;
      } else {
        speeds[c] = SUFFIX(collide_cell)(t, o, omega);
      }
      for (int k = 0; k < 9; k++) dst[(size_t)k * n + c] = o[k];
    }
  }

  /* sequential row-major accumulation (the reduction ORDER is part of the golden pin) */
#if VARIANT == VARIANT_REF32
  float tot = 0;
  for (size_t c = 0; c < n; c++) tot += speeds[c];          /* d2q9-bgk.c:416-420 */
  return (double)(tot / (float)tot_cells);                  /* :423 */
#else
  double tot = 0.0;
  for (size_t c = 0; c < n; c++) if (!obst[c]) tot += (double)speeds[c];
  return tot / (double)tot_cells;
#endif
}

/* av_velocity on a resident state, d2q9-bgk.c:426-475 (feeds the Reynolds number) */
double SUFFIX(canon_av_velocity)(int nx, int ny, const int* obst, const REAL* f)
{
  const size_t n = (size_t)nx * (size_t)ny;
  long cnt = 0;
  REAL tot = 0.0;
  for (size_t c = 0; c < n; c++) {
    if (obst[c]) continue;
    REAL ld = 0.0;
    for (int k = 0; k < 9; k++) ld += f[(size_t)k * n + c];
    const REAL ux = (f[1 * n + c] + f[5 * n + c] + f[8 * n + c]
                     - (f[3 * n + c] + f[6 * n + c] + f[7 * n + c])) / ld;
    const REAL uy = (f[2 * n + c] + f[5 * n + c] + f[6 * n + c]
                     - (f[4 * n + c] + f[7 * n + c] + f[8 * n + c])) / ld;
    tot += sqrt((ux * ux) + (uy * uy));
    cnt++;
  }
  return (double)(tot / (REAL)cnt);
}

/* macroscopic fields of the final state, d2q9-bgk.c:857-897: ux, uy, |u|, pressure */
void SUFFIX(canon_macroscopic)(int nx, int ny, REAL density, const int* obst, const REAL* f,
                               REAL* out_ux, REAL* out_uy, REAL* out_u, REAL* out_p)
{
  const size_t n = (size_t)nx * (size_t)ny;
  const REAL c_sq = 1.0 / 3.0;
  for (size_t c = 0; c < n; c++) {
    if (obst[c]) {
      out_ux[c] = out_uy[c] = out_u[c] = 0.0;
      out_p[c] = density * c_sq;
      continue;
    }
    REAL ld = 0.0;
    for (int k = 0; k < 9; k++) ld += f[(size_t)k * n + c];
    const REAL ux = (f[1 * n + c] + f[5 * n + c] + f[8 * n + c]
                     - (f[3 * n + c] + f[6 * n + c] + f[7 * n + c])) / ld;
    const REAL uy = (f[2 * n + c] + f[5 * n + c] + f[6 * n + c]
                     - (f[4 * n + c] + f[7 * n + c] + f[8 * n + c])) / ld;
    out_ux[c] = ux; out_uy[c] = uy;
    out_u[c] = sqrt((ux * ux) + (uy * uy));
    out_p[c] = ld * c_sq;
  }
}

/* run `iters` steps from `f` (ping-pong with scratch), av_vels[iters] as double.
 * On return f holds the final state.  Returns 0, or -1 on allocation failure. */
int SUFFIX(canon_run)(int nx, int ny, REAL density, REAL accel, REAL omega,
                      const int* obst, REAL* f, int iters, double* av_vels)
{
  const size_t n = (size_t)nx * (size_t)ny;
  REAL* tmp = (REAL*)malloc(sizeof(REAL) * 9 * n);
  REAL* speeds = (REAL*)malloc(sizeof(REAL) * n);
  if (!tmp || !speeds) { free(tmp); free(speeds); return -1; }
  long tot_cells = 0;
  for (size_t c = 0; c < n; c++) tot_cells += !obst[c];
  REAL* a = f; REAL* b = tmp;
  for (int t = 0; t < iters; t++) {
    const double av = SUFFIX(canon_step)(nx, ny, density, accel, omega, obst, a, b, speeds,
                                         tot_cells, 1);
    if (av_vels) av_vels[t] = av;
    REAL* s = a; a = b; b = s;
  }
  if (a != f) memcpy(f, a, sizeof(REAL) * 9 * n);
  free(tmp); free(speeds);
  return 0;
}
