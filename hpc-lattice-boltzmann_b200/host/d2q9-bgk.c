/*
 * d2q9-bgk.c -- host program of the B200-native D2Q9-BGK lattice-Boltzmann solver (plain C99).
 *
 * Drop-in for the reference executable of AlexDalt/HPC-Lattice-Boltzmann: same command line
 *     d2q9-bgk.exe <paramfile> <obstaclefile>
 * same input formats (d2q9-bgk.c:499-525 params, :615-628 obstacles), same diagnostics, same
 * stdout block (:283-287) and the same av_vels.dat / final_state.dat formats (:900, :915), so the
 * reference's check/check.py accepts the outputs unchanged.  Everything between reading the
 * inputs and writing the outputs -- the reference's timestep(), accelerate_flow(), comp_func()
 * and av_velocity() -- runs on the GPU behind the C ABI in include/lbm_b200.h.  There is no CPU
 * fallback: without a B200-class GPU the program reports the library's error and exits.
 *
 * Extra knobs are environment variables only (the positional interface is unchanged):
 *   LBM_GPUS=<n>            row-slab the lattice over n GPUs of this node (default 1)
 *   LBM_SKIP_FINAL_STATE=1  do not write final_state.dat (synthetic multi-GB cases)
 *   LBM_ITERS=<n>           override maxIters from the params file (benchmarks)
 *   LBM_SERIAL_PARSE=1      read the obstacle file with the reference's fscanf loop only
 *   LBM_PARSE_ONLY=1        read both input files, print the blocked-cell count and a checksum of the
 *                           obstacle map, exit (no GPU needed: used by the CPU tests of the parser)
 */
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "lbm_b200.h"

#define NSPEEDS 9
#define FINAL_STATE_FILE "final_state.dat"
#define AV_VELS_FILE     "av_vels.dat"

typedef struct {
  int   nx, ny;         /* grid */
  int   max_iters;      /* timesteps */
  int   reynolds_dim;   /* length scale for the Reynolds number */
  float density, accel, omega;
} run_config;

/* same message shapes as the reference's die()/usage() (d2q9-bgk.c:933-945) */
static void die_at(const char* message, int line, const char* file)
{
  fprintf(stderr, "Error at line %d of file %s:\n", line, file);
  fprintf(stderr, "%s\n", message);
  fflush(stderr);
  exit(EXIT_FAILURE);
}
#define DIE(msg) die_at((msg), __LINE__, __FILE__)

static void gpu_check(int rc, const char* what, int line)
{
  if (rc == 0) return;
  fprintf(stderr, "GPU engine error during '%s' on line %d: %s\n", what, line, lbm_last_error());
  fflush(stderr);
  exit(EXIT_FAILURE);
}
#define GPU(call) gpu_check((call), #call, __LINE__)

static double wall_seconds(void)
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + tv.tv_usec / 1000000.0;
}

static void read_param_file(const char* path, run_config* cfg)
{
  char message[1100];
  FILE* fp = fopen(path, "r");
  if (fp == NULL) {
    snprintf(message, sizeof message, "could not open input parameter file: %s", path);
    DIE(message);
  }
  /* one scalar per line, in this order; a failed conversion names the field like the reference */
  struct { const char* fmt; void* dst; const char* name; } fields[] = {
    {"%d\n", &cfg->nx, "nx"},
    {"%d\n", &cfg->ny, "ny"},
    {"%d\n", &cfg->max_iters, "maxIters"},
    {"%d\n", &cfg->reynolds_dim, "reynolds_dim"},
    {"%f\n", &cfg->density, "density"},
    {"%f\n", &cfg->accel, "accel"},
    {"%f\n", &cfg->omega, "omega"},
  };
  for (size_t i = 0; i < sizeof fields / sizeof fields[0]; i++) {
    if (fscanf(fp, fields[i].fmt, fields[i].dst) != 1) {
      snprintf(message, sizeof message, "could not read param file: %s", fields[i].name);
      DIE(message);
    }
  }
  fclose(fp);
}

/* The reference reads the obstacle file with fscanf("%d %d %d\n") in a serial loop
 * (d2q9-bgk.c:615-628); at 16384 x 16384 that file has ~4 M lines.  Fast path: the file is mapped,
 * cut into one byte range per thread at line boundaries, and every thread parses its lines with a
 * hand-written integer scanner.  It accepts exactly the well-formed shape "x y b" (blanks between
 * the numbers, optional blanks / CR before the newline, empty lines skipped, last line may lack
 * the newline).  Anything else -- a malformed line, a value out of range, b != 1 -- makes it
 * report "irregular", and the caller re-reads the whole file with the reference's own fscanf
 * loop, which then produces the reference's diagnostics for the first offending entry. */
static const char* scan_int(const char* p, const char* end, long* out)
{
  int neg = 0;
  if (p < end && (*p == '-' || *p == '+')) neg = *p++ == '-';
  if (p >= end || *p < '0' || *p > '9') return NULL;
  long v = 0;
  while (p < end && *p >= '0' && *p <= '9') {
    v = v * 10 + (*p++ - '0');
    if (v > 0x7fffffffL) return NULL;
  }
  *out = neg ? -v : v;
  return p;
}

static int parse_obstacles_fast(const char* path, const run_config* cfg, int* blocked_map)
{
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return 0;
  struct stat sb;
  if (fstat(fd, &sb) != 0 || sb.st_size == 0) { close(fd); return sb.st_size == 0 ? 1 : 0; }
  const size_t n = (size_t)sb.st_size;
  const char* text = mmap(NULL, n, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (text == MAP_FAILED) return 0;
  int irregular = 0;
  const long nx = cfg->nx, ny = cfg->ny;
#pragma omp parallel reduction(| : irregular)
  {
#ifdef _OPENMP
    const size_t t = (size_t)omp_get_thread_num(), nt = (size_t)omp_get_num_threads();
#else
    const size_t t = 0, nt = 1;
#endif
    /* my lines: those that START in [t n / nt, (t+1) n / nt); both ends moved up to a line start */
    size_t lo = n * t / nt, hi = n * (t + 1) / nt;
    while (lo > 0 && lo < n && text[lo - 1] != '\n') lo++;
    while (hi > 0 && hi < n && text[hi - 1] != '\n') hi++;
    const char* p = text + lo;
    const char* end = text + hi;
    while (p < end && !irregular) {
      while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) p++;
      if (p >= end) break;
      long x, y, b;
      const char* q = scan_int(p, end, &x);
      if (q == NULL || q >= end || (*q != ' ' && *q != '\t')) { irregular = 1; break; }
      while (q < end && (*q == ' ' || *q == '\t')) q++;
      q = scan_int(q, end, &y);
      if (q == NULL || q >= end || (*q != ' ' && *q != '\t')) { irregular = 1; break; }
      while (q < end && (*q == ' ' || *q == '\t')) q++;
      q = scan_int(q, end, &b);
      if (q == NULL) { irregular = 1; break; }
      while (q < end && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
      if (q < end && *q != '\n') { irregular = 1; break; }
      if (x < 0 || x > nx - 1 || y < 0 || y > ny - 1 || b != 1) { irregular = 1; break; }
      blocked_map[(size_t)y * (size_t)nx + (size_t)x] = 1;     /* duplicates: every writer stores 1 */
      p = q;
    }
  }
  munmap((void*)text, n);
  return !irregular;
}

static int* read_obstacle_file(const char* path, const run_config* cfg)
{
  char message[1100];
  const size_t cells = (size_t)cfg->nx * cfg->ny;
  int* blocked_map = calloc(cells, sizeof(int));
  if (blocked_map == NULL) DIE("cannot allocate column memory for obstacles");
  FILE* fp = fopen(path, "r");
  if (fp == NULL) {
    snprintf(message, sizeof message, "could not open input obstacles file: %s", path);
    DIE(message);
  }
  if (getenv("LBM_SERIAL_PARSE") == NULL && parse_obstacles_fast(path, cfg, blocked_map)) {
    fclose(fp);
    return blocked_map;
  }
  /* the reference's loop (d2q9-bgk.c:615-628): exact semantics and diagnostics for odd inputs */
  memset(blocked_map, 0, cells * sizeof(int));
  int x, y, blocked, got;
  while ((got = fscanf(fp, "%d %d %d\n", &x, &y, &blocked)) != EOF) {
    if (got != 3) DIE("expected 3 values per line in obstacle file");
    if (x < 0 || x > cfg->nx - 1) DIE("obstacle x-coord out of range");
    if (y < 0 || y > cfg->ny - 1) DIE("obstacle y-coord out of range");
    if (blocked != 1) DIE("obstacle blocked value should be 1");
    blocked_map[(size_t)y * cfg->nx + x] = blocked;
  }
  fclose(fp);
  return blocked_map;
}

/* final_state.dat, the reference's row-major order and format string (d2q9-bgk.c:857-900).  Rows
 * are formatted in parallel (each thread snprintf()s whole rows into its own buffer -- the same
 * libc conversion, hence byte-identical text) and written out in order, batch by batch: at
 * 16384 x 16384 the file has 268 M lines and a serial fprintf loop would take minutes. */
static void write_final_state(const run_config* cfg, const int* blocked_map, const float* ux,
                              const float* uy, const float* speed, const float* pressure)
{
  FILE* fp = fopen(FINAL_STATE_FILE, "w");
  if (fp == NULL) DIE("could not open file output file");
  const int nx = cfg->nx, ny = cfg->ny;
  const size_t line_cap = 128;                    /* 2 ints + 4 x "%.12E" + flag: < 110 bytes */
  int batch = 64;
  if (batch > ny) batch = ny;
  char* text = malloc((size_t)batch * nx * line_cap);
  size_t* used = malloc(sizeof(size_t) * (size_t)batch);
  if (text == NULL || used == NULL) DIE("cannot allocate memory for output buffers");
  for (int y0 = 0; y0 < ny; y0 += batch) {
    const int rows = (ny - y0 < batch) ? ny - y0 : batch;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; r++) {
      const int y = y0 + r;
      char* out = text + (size_t)r * nx * line_cap;
      size_t n = 0;
      for (int x = 0; x < nx; x++) {
        const size_t c = (size_t)y * nx + x;
        n += (size_t)snprintf(out + n, line_cap, "%d %d %.12E %.12E %.12E %.12E %d\n", x, y, ux[c],
                              uy[c], speed[c], pressure[c], blocked_map[c]);
      }
      used[r] = n;
    }
    for (int r = 0; r < rows; r++)
      if (fwrite(text + (size_t)r * nx * line_cap, 1, used[r], fp) != used[r])
        DIE("could not write to output file");
  }
  free(used);
  free(text);
  fclose(fp);
}

static void write_av_vels(const run_config* cfg, const float* av_vels)
{
  FILE* fp = fopen(AV_VELS_FILE, "w");
  if (fp == NULL) DIE("could not open file output file");
  for (int t = 0; t < cfg->max_iters; t++) fprintf(fp, "%d:\t%.12E\n", t, av_vels[t]);
  fclose(fp);
}

static int env_flag(const char* name, int dflt)
{
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

int main(int argc, char* argv[])
{
  if (argc != 3) {
    fprintf(stderr, "Usage: %s <paramfile> <obstaclefile>\n", argv[0]);
    exit(EXIT_FAILURE);
  }

  run_config cfg;
  const double parse_tic = wall_seconds();
  read_param_file(argv[1], &cfg);
  int* blocked_map = read_obstacle_file(argv[2], &cfg);
  const double parse_toc = wall_seconds();
  if (env_flag("LBM_PARSE_ONLY", 0)) {
    unsigned long long blocked = 0, sum = 1469598103934665603ULL;
    for (size_t i = 0; i < (size_t)cfg.nx * cfg.ny; i++) {
      blocked += blocked_map[i] != 0;
      sum = (sum ^ (unsigned long long)(blocked_map[i] != 0)) * 1099511628211ULL;      /* FNV-1a */
    }
    printf("parsed: %d x %d, blocked=%llu, checksum=%016llx\n", cfg.nx, cfg.ny, blocked, sum);
    free(blocked_map);
    return EXIT_SUCCESS;
  }
  cfg.max_iters = env_flag("LBM_ITERS", cfg.max_iters);
  const int ngpus = env_flag("LBM_GPUS", 1);
  const int skip_final_state = env_flag("LBM_SKIP_FINAL_STATE", 0);
  const size_t cells = (size_t)cfg.nx * cfg.ny;

  float* av_vels = malloc(sizeof(float) * (size_t)(cfg.max_iters > 0 ? cfg.max_iters : 1));
  if (av_vels == NULL) DIE("cannot allocate memory for cells");

  lbm_lattice* lattice = NULL;
  lbm_params params = {cfg.nx, cfg.ny, cfg.density, cfg.accel, cfg.omega};
  GPU(lbm_create(&lattice, &params, blocked_map, ngpus));
  /* ux, uy, |u|, pressure: page-locked, so the four result planes come back at full PCIe speed */
  void* pinned = NULL;
  GPU(lbm_host_alloc(&pinned, (unsigned long long)(sizeof(float) * 4 * cells)));
  float* fields = pinned;

  /* timed region, as in the reference (d2q9-bgk.c:155 .. :275): state generation/transfer,
   * all timesteps, results back on the host */
  const double tic = wall_seconds();
  GPU(lbm_init_equilibrium(lattice));
  GPU(lbm_run(lattice, cfg.max_iters, av_vels));
  const double loop_ms = lbm_last_run_ms(lattice);
  float final_av = 0.0f;
  GPU(lbm_av_velocity(lattice, &final_av));
  GPU(lbm_macroscopic(lattice, fields, fields + cells, fields + 2 * cells, fields + 3 * cells));
  const double toc = wall_seconds();

  struct rusage ru;
  getrusage(RUSAGE_SELF, &ru);
  const double usrtim = ru.ru_utime.tv_sec + ru.ru_utime.tv_usec / 1000000.0;
  const double systim = ru.ru_stime.tv_sec + ru.ru_stime.tv_usec / 1000000.0;

  /* calc_reynolds, d2q9-bgk.c:815-820 */
  const float viscosity = 1.0 / 6.0 * (2.0 / cfg.omega - 1.0);
  const float reynolds = final_av * cfg.reynolds_dim / viscosity;

  printf("==done==\n");
  printf("Reynolds number:\t\t%.12E\n", reynolds);
  printf("Elapsed time:\t\t\t%.6lf (s)\n", toc - tic);
  printf("Elapsed user CPU time:\t\t%.6lf (s)\n", usrtim);
  printf("Elapsed system CPU time:\t%.6lf (s)\n", systim);
  /* additions (after the reference's block, so scripts that parse it keep working) */
  const double lups = (double)cells * cfg.max_iters;
  printf("Input parse:\t\t\t%.6lf (s)\n", parse_toc - parse_tic);
  printf("GPU timestep loop:\t\t%.6lf (s)\n", loop_ms / 1e3);
  if (loop_ms > 0.0) {
    const double mlups = lups / (loop_ms / 1e3) / 1e6;
    printf("MLUPS (timestep loop):\t\t%.1f\n", mlups);
    printf("HBM traffic (72 B/update):\t%.1f GB/s\n", mlups * 72.0 / 1e3);
  }
  if (toc > tic) printf("MLUPS (elapsed time):\t\t%.1f\n", lups / (toc - tic) / 1e6);
  printf("GPUs:\t\t\t\t%d (%s)\n", ngpus, lbm_config_string(lattice));

  if (!skip_final_state)
    write_final_state(&cfg, blocked_map, fields, fields + cells, fields + 2 * cells, fields + 3 * cells);
  write_av_vels(&cfg, av_vels);

  lbm_destroy(lattice);
  lbm_host_free(fields);
  free(av_vels);
  free(blocked_map);
  return EXIT_SUCCESS;
}
