/*
 * spectralbte_b200/host/boltz_b200.c -- C host driver on top of libsbte_b200.so (device-resident path).
 *
 *     boltz_b200 <input file> <output-flags file>        (both looked up under ./input/)
 *
 * Mirrors the reference driver /root/reference/exec/boltz.c for one rank: same command line, same
 * input keywords (src/input.c:30-116, defaults :132-183), same mesh file (src/mesh_setup.c:27-71), same
 * Weights/N%d_isotropic_L_v%g_lambda%g.wts naming/format (src/weights.c:66-124: load if present unless
 * Recompute_weights, else generate and store), same Data/moments_<input> text output
 * (src/output.c:100-135,159-213,262-343).  The distribution function lives on the GPU for the whole
 * run; only moments (and the N-point slice in 0D) come back at output steps.
 * Scope: one species ("default"), isotropic weights, Init_field 0/2/4/5 (0D) and 0/3/6 (1D).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sbte_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define CHECK(call)                                                        \
  do {                                                                     \
    if ((call) != 0) {                                                     \
      printf("boltz_b200: %s failed: %s\n", #call, sbte_last_error());     \
      exit(1);                                                             \
    }                                                                      \
  } while (0)

typedef struct {
  int N, nT, order, dataFreq, restart, initFlag, homogFlag, weightFlag, isoFlag, num_species;
  double L_v, Kn, lambda, dt, restart_time;
  char meshFile[80];
} params;

typedef struct { int dens, vel, temp, pres, marginal, slice, ent; } outflags;

/* ---- keyword / next-token reader, as src/input.c:201-235 ---- */
static int next_token(FILE *fp, char *tok) { return fscanf(fp, "%79s", tok) == 1; }
static void need(int ok) {
  if (!ok) {
    printf("\nError while reading the input file!\nPlease, change the input file...\n");
    exit(0);
  }
}
static long read_int(FILE *fp) { unsigned long v; need(fscanf(fp, "%lu", &v) == 1); return (long)v; }
static double read_double(FILE *fp) { double v; need(fscanf(fp, "%lf", &v) == 1); return v; }

static void read_input(const char *name, params *p) {
  char path[200], tok[80] = "dummy";
  FILE *fp;
  /* defaults: src/input.c:132-183 (Init_field has none) */
  p->homogFlag = 0; p->weightFlag = 0; p->N = 16; p->L_v = 5.; p->Kn = 1; p->lambda = 1; p->order = 1;
  p->restart = 0; p->restart_time = 85500; p->dt = -1.0; p->nT = 1000; p->dataFreq = 10; p->isoFlag = 0;
  p->num_species = 1; p->initFlag = -1; strcpy(p->meshFile, "not set");
  snprintf(path, sizeof(path), "./input/%s", name);
  printf("Opening input file %s\n", path);
  fp = fopen(path, "r");
  if (!fp) { printf("Error - input file not found\n"); exit(1); }
  while (strcmp(tok, "Stop") != 0) {
    need(next_token(fp, tok));
    if (!strcmp(tok, "N")) p->N = (int)read_int(fp);
    else if (!strcmp(tok, "L_v")) p->L_v = read_double(fp);
    else if (!strcmp(tok, "Knudsen")) p->Kn = read_double(fp);
    else if (!strcmp(tok, "Lambda")) p->lambda = read_double(fp);
    else if (!strcmp(tok, "Time_step")) p->dt = read_double(fp);
    else if (!strcmp(tok, "Number_of_time_steps")) p->nT = (int)read_int(fp);
    else if (!strcmp(tok, "Space_order")) p->order = (int)read_int(fp);
    else if (!strcmp(tok, "Data_writing_frequency")) p->dataFreq = (int)read_int(fp);
    else if (!strcmp(tok, "Restart")) p->restart = (int)read_int(fp);
    else if (!strcmp(tok, "Restart_time")) p->restart_time = read_double(fp);
    else if (!strcmp(tok, "Init_field")) p->initFlag = (int)read_int(fp);
    else if (!strcmp(tok, "Bound_cond")) (void)read_int(fp);
    else if (!strcmp(tok, "SpaceInhom")) p->homogFlag = (int)read_int(fp);
    else if (!strcmp(tok, "Recompute_weights")) p->weightFlag = (int)read_int(fp);
    else if (!strcmp(tok, "Anisotropic")) p->isoFlag = (int)read_int(fp);
    else if (!strcmp(tok, "mesh_file")) need(next_token(fp, p->meshFile));
    else if (!strcmp(tok, "num_species")) {
      int i;
      p->num_species = (int)read_int(fp);
      for (i = 0; i < p->num_species; i++) {
        need(next_token(fp, tok));
        if (strcmp(tok, "default") != 0) { printf("boltz_b200: only the default species is supported (%s)\n", tok); exit(1); }
      }
      strcpy(tok, "dummy");
    }
  }
  fclose(fp);
  if (!strcmp(p->meshFile, "not set") && p->homogFlag == 1) { printf("Error: please specify the mesh\n"); exit(1); }
  if (p->num_species != 1 || p->isoFlag != 0) { printf("boltz_b200: one species, isotropic weights only\n"); exit(1); }
  if (p->restart && p->homogFlag == 0) { printf("boltz_b200: Restart applies to the inhomogeneous case only\n"); exit(1); }
  printf("done with input file\n");
}

/* src/output.c:417-443 */
static void read_flags(const char *name, outflags *o) {
  char path[200], tok[80] = "dummy";
  FILE *fp;
  o->dens = o->vel = o->temp = o->pres = o->marginal = o->slice = o->ent = 1;
  snprintf(path, sizeof(path), "./input/%s", name);
  fp = fopen(path, "r");
  if (!fp) { printf("Error - output flags file not found\n"); exit(1); }
  need(next_token(fp, tok));
  while (strcmp(tok, "Stop") != 0) {
    need(next_token(fp, tok));
    if (!strcmp(tok, "density")) o->dens = (int)read_int(fp);
    else if (!strcmp(tok, "velocity")) o->vel = (int)read_int(fp);
    else if (!strcmp(tok, "temperature")) o->temp = (int)read_int(fp);
    else if (!strcmp(tok, "pressure")) o->pres = (int)read_int(fp);
    else if (!strcmp(tok, "marginal")) o->marginal = (int)read_int(fp);
    else if (!strcmp(tok, "slice")) o->slice = (int)read_int(fp);
    if (!strcmp(tok, "entropy")) o->ent = (int)read_int(fp);
  }
  fclose(fp);
}

/* src/mesh_setup.c:27-71,121-146 (+ right ghosts by extension, :165-175) */
static void make_mesh(const char *name, int order, int *nX, double **x, double **dx) {
  char path[300], line[300];
  FILE *fp;
  int zones, z, j, cnt, total;
  double edge = 0.0;
  snprintf(path, sizeof(path), "./input/%s", name);
  printf("Opening %s\n", path);
  fp = fopen(path, "r");
  if (!fp) { printf("Error - mesh file not found\n"); exit(1); }
  need(fscanf(fp, "%299s", line) == 1);
  need(fscanf(fp, "%d", nX) == 1);
  need(fscanf(fp, "%d", &zones) == 1);
  if (zones < 1) { printf("Error - bad number of zones listed in mesh generation\n"); exit(0); }
  total = *nX + 2 * order;
  *x = malloc(sizeof(double) * total);
  *dx = malloc(sizeof(double) * total);
  cnt = order;
  for (z = 0; z < zones; z++) {
    int nz; double Lz, d;
    need(fscanf(fp, "%d", &nz) == 1);
    need(fscanf(fp, "%lf", &Lz) == 1);
    d = Lz / (double)nz;
    for (j = 0; j < nz && cnt < *nX + order; j++) { (*dx)[cnt] = d; (*x)[cnt] = edge + 0.5 * d; edge += d; cnt++; }
  }
  fclose(fp);
  for (j = order - 1; j >= 0; j--) { (*dx)[j] = (*dx)[j + 1]; (*x)[j] = (*x)[j + 1] - (*dx)[j + 1]; }
  for (j = *nX + order; j < total; j++) { (*dx)[j] = (*dx)[j - 1]; (*x)[j] = (*x)[j - 1] + (*dx)[j - 1]; }
  printf("Loaded mesh!\n");
}

/* grids: src/initializer.c:66-82 (0D) / :256-266 (1D) */
static void make_grids(const params *p, double *v, double *eta) {
  const int N = p->N;
  const double dv = 2 * p->L_v / (N - 1);
  double deta, L_eta;
  int i;
  for (i = 0; i < N; i++) v[i] = -p->L_v + i * dv;
  if (p->homogFlag == 0) {
    deta = (2 * M_PI / N) / dv;
    L_eta = ((N % 2) == 0) ? 0.5 * N * deta : 0.5 * (N - 1) * deta;
  } else {
    L_eta = 0.5 * (N - 1) * M_PI / p->L_v;
    deta = M_PI * (N - 1) / (N * p->L_v);
  }
  for (i = 0; i < N; i++) eta[i] = -L_eta + i * deta;
}

/* src/initializer.c:91-198 */
static void init_hom(const params *p, const double *v, double *f) {
  const int N = p->N;
  int i, j, k;
  printf("Initializing...%d\n", p->initFlag);
  for (i = 0; i < N; i++)
    for (j = 0; j < N; j++)
      for (k = 0; k < N; k++) {
        const double r2 = v[i] * v[i] + v[j] * v[j] + v[k] * v[k];
        double val;
        switch (p->initFlag) {
          case 0: { const double sigma = 0.3 * p->L_v, S = 10.0;
                    val = exp(-1 * S * (sqrt(r2) - sigma) * (sqrt(r2) - sigma) / (sigma * sigma)) / (S * S); break; }
          case 2: { const double K = 1 - exp(-5.5 / 6.0), T = 1.0;
                    val = (exp(-r2 / (2 * K * T * T))) / (2.0 * pow(2 * M_PI * K * T * T, 1.5)) *
                          ((5 * K - 3) / K + (1 - K) * r2 / (K * K * T * T)); break; }
          case 4: val = pow(0.5 / M_PI, 1.5) * exp(-0.5 * r2); break;
          case 5: val = (1 + 0.1 * sin(r2)) * exp(-r2) / (M_PI * sqrt(M_PI)); break;
          default: printf("boltz_b200: Init_field %d not implemented for the homogeneous case\n", p->initFlag); exit(1);
        }
        f[k + N * (j + N * i)] = val;
      }
}

/* src/initializer.c:297-421 */
static void init_inhom(const params *p, const double *v, int nX, double *slab) {
  const int N = p->N, order = p->order;
  const long n3 = (long)N * N * N;
  double rho_l = 1, ux_l = 0, T_l = 1, rho_r = 1, ux_r = 0, T_r = 1;
  int i, j, k, l;
  switch (p->initFlag) {
    case 0: { const double Ma = 1;
              rho_l = 4.0 * Ma * Ma / (Ma * Ma + 3.0); T_l = (5.0 * Ma * Ma - 1.0) * (Ma * Ma + 3.0) / (16.0 * Ma * Ma); break; }
    case 1: T_r = 1.0; break;                                  /* sudden heating: wall at 2*TWall */
    case 2: T_r = 2.0; ux_l = -1.0; ux_r = -1.0; break;        /* shifted Maxwellian, uses the "left" state everywhere */
    case 3: T_r = 1.5; break;
    case 5: break;                                             /* Poiseuille: rho 1, T 1 at rest (the "left" state) */
    case 6: ux_l = 1.2972; rho_r = 1.297; ux_r = 1.0; T_r = 1.195; break;
    default: printf("boltz_b200: Init_field %d not implemented for the inhomogeneous case\n", p->initFlag); exit(1);
  }
  memset(slab, 0, sizeof(double) * (size_t)(nX + 2 * order) * n3);
  for (l = order; l < nX + order; l++) {
    const int left = (p->initFlag == 3 || p->initFlag == 1) ? 0 : ((p->initFlag == 2 || p->initFlag == 5) ? 1 : (l < nX / 2));
    const double rho = left ? rho_l : rho_r, T = left ? T_l : T_r;
    const double ux = (p->initFlag == 6 || p->initFlag == 2) ? (left ? ux_l : ux_r) : 0.0;
    for (i = 0; i < N; i++)
      for (j = 0; j < N; j++)
        for (k = 0; k < N; k++)
          slab[l * n3 + k + N * (j + N * i)] =
              (p->initFlag == 1)   /* src/initializer.c:381 uses exp(-v^2 / 2T) for this case */
                  ? pow(0.5 / (M_PI * T), 1.5) * exp(-(0.5 / T) * (v[i] * v[i] + v[j] * v[j] + v[k] * v[k]))
                  : rho * exp(-((v[i] - ux) * (v[i] - ux) + v[j] * v[j] + v[k] * v[k]) / T) / ((T * M_PI) * sqrt(T * M_PI));
  }
}

/* src/weights.c:66-124: load the stored file unless Recompute_weights, else generate (on the device) and store */
/* primary = 0: a further GPU of the same run; the file exists by now (loaded or just written by the first GPU) */
static void setup_weights(sbte_ctx *ctx, const params *p, int primary) {
  char name[200];
  FILE *fp;
  snprintf(name, sizeof(name), "Weights/N%d_isotropic_L_v%g_lambda%g.wts", p->N, p->L_v, p->lambda);
  if (!primary) { CHECK(sbte_weights_load_file(ctx, name)); return; }
  if (p->weightFlag == 0 && (fp = fopen(name, "r"))) {
    fclose(fp);
    printf("Loading weights from file %s\n", name);
    CHECK(sbte_weights_load_file(ctx, name));
    return;
  }
  if (p->weightFlag == 0) printf("Stored weights not found for this configuration, generating ...\n");
  else printf("Fresh version of weights being computed and stored for this configuration\n");
  CHECK(sbte_weights_generate_iso(ctx, p->lambda));
  CHECK(sbte_weights_save_file(ctx, name));
}

/* ---- restart files, src/restart.c:24-109: Restart/<input>_rank0_default.plt = the owned cells as raw
 * doubles, Restart/<input>_time.plt = the step counter (an int). ---- */
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void store_restart(const char *input, const double *slab_h, int nX, int order, long n3, int t, int rank) {
  char name[300];
  FILE *fp;
  if (rank == 0) {
    snprintf(name, sizeof(name), "Restart/%s_time.plt", input);
    fp = fopen(name, "w");
    if (!fp) { printf("Something happened when trying to save the time\n"); exit(1); }
    printf("Time stored: %d\n", t);
    fwrite(&t, sizeof(int), 1, fp);
    fclose(fp);
  }
  snprintf(name, sizeof(name), "Restart/%s_rank%d_%s.plt", input, rank, "default");
  printf("%s\n", name);
  fp = fopen(name, "w");
  if (!fp || fwrite(slab_h + (size_t)order * n3, sizeof(double), (size_t)nX * n3, fp) != (size_t)nX * n3) {
    printf("Something happened when trying to save the pdf\n");
    exit(1);
  }
  fclose(fp);
}

static void load_restart(const char *input, double *slab_h, int nX, int order, long n3, int *t, int rank) {
  char name[300];
  FILE *fp;
  snprintf(name, sizeof(name), "Restart/%s_rank%d_%s.plt", input, rank, "default");
  printf("Loading data %s\n", name);
  fp = fopen(name, "r");
  if (!fp) { printf("Error: unable to open restarted file %s\n", name); exit(1); }
  if (fread(slab_h + (size_t)order * n3, sizeof(double), (size_t)nX * n3, fp) != (size_t)nX * n3) {
    printf("Error reloading pdf file\n");
    exit(1);
  }
  fclose(fp);
  snprintf(name, sizeof(name), "Restart/%s_time.plt", input);
  fp = fopen(name, "r");
  if (!fp || fread(t, sizeof(int), 1, fp) != 1) { printf("Error: unable to read %s\n", name); exit(1); }
  fclose(fp);
  printf("t: %d\n", *t);
}

/* ---- one host thread per GPU for a run of steps: the GPUs wait for each other on the device (peer-memory halo),
 * so no GPU's step may be held back by a host-side call made on behalf of another one ---- */
typedef struct {
  sbte_slab *slab;
  sbte_ctx *ctx;
  double Kn;
  int nsteps, rc;
  char err[256];
} step_job;

static void *step_worker(void *arg) {
  step_job *j = arg;
  int i;
  j->rc = 0;
  for (i = 0; i < j->nsteps && !j->rc; i++) j->rc = sbte_slab_step(j->slab, j->Kn, SBTE_K2_AUTO);
  if (!j->rc) j->rc = sbte_sync(j->ctx);
  if (j->rc) snprintf(j->err, sizeof(j->err), "%s", sbte_last_error());
  return NULL;
}

static void run_steps(sbte_slab **slabs, sbte_ctx **ctxs, int G, double Kn, int nsteps) {
  step_job jobs[64];
  pthread_t th[64];
  int r;
  for (r = 0; r < G; r++) { jobs[r].slab = slabs[r]; jobs[r].ctx = ctxs[r]; jobs[r].Kn = Kn; jobs[r].nsteps = nsteps; }
  if (G == 1) step_worker(&jobs[0]);
  else {
    for (r = 0; r < G; r++) pthread_create(&th[r], NULL, step_worker, &jobs[r]);
    for (r = 0; r < G; r++) pthread_join(th[r], NULL);
  }
  for (r = 0; r < G; r++)
    if (jobs[r].rc) { printf("boltz_b200: time step failed on GPU %d: %s\n", r, jobs[r].err); exit(1); }
}

int main(int argc, char **argv) {
  params p;
  outflags of;
  double *v, *eta;
  sbte_ctx *ctx;
  FILE *out;
  char outname[300];
  long n3;
  int t, l, outputCount = 0;
  if (argc < 3) { printf("usage: boltz_b200 <input file> <output flags file>\n"); return 1; }
  /* several GPUs in one process wait for each other on the device: no kernel may be loaded lazily behind such a wait */
  if (getenv("SBTE_GPUS") && atoi(getenv("SBTE_GPUS")) > 1) setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  read_input(argv[1], &p);
  read_flags(argv[2], &of);
  n3 = (long)p.N * p.N * p.N;
  v = malloc(sizeof(double) * p.N);
  eta = malloc(sizeof(double) * p.N);
  make_grids(&p, v, eta);
  CHECK(sbte_create(&ctx, p.N, p.L_v, v, eta, getenv("SBTE_DEVICE") ? atoi(getenv("SBTE_DEVICE")) : 0));
  printf("Initializing weight info\n");
  setup_weights(ctx, &p, 1);
  snprintf(outname, sizeof(outname), "Data/moments_%s", argv[1]);
  printf("Opening output files \n");
  out = fopen(outname, p.restart ? "a" : "w");   /* src/output.c:256-262 */
  if (!out) { printf("boltz_b200: cannot open %s\n", outname); return 1; }
  printf("Done with all setup, starting main loop\n");

  if (p.homogFlag == 0) {
    /* ---------------- space homogeneous: exec/boltz.c:186-250 ---------------- */
    double *f = malloc(sizeof(double) * n3), mom[8], *d_f, *d_mom;
    init_hom(&p, v, f);
    CHECK(sbte_dev_alloc((void **)&d_f, sizeof(double) * n3));
    CHECK(sbte_dev_alloc((void **)&d_mom, sizeof(double) * 8));
    CHECK(sbte_h2d(ctx, d_f, f, sizeof(double) * n3));
    fprintf(out, "#Time ");
    if (of.dens) fprintf(out, "Density ");
    if (of.vel) fprintf(out, "Velocity_x ");
    if (of.temp) fprintf(out, "Temperature ");
    if (of.pres) fprintf(out, "Pressure ");
    if (of.ent) fprintf(out, "Pos/neg_energy_ratio ");
    if (of.slice) for (l = 0; l < p.N; l++) fprintf(out, "v:%le ", v[l]);
    fprintf(out, "\n");
    for (t = 0; t <= p.nT; t++) {
      if (t > 0) {
        printf("In step %d of %d\n", t, p.nT);
        CHECK(sbte_step_0d(ctx, d_f, p.dt, p.Kn, p.order, SBTE_K2_AUTO));
        outputCount++;
      }
      if (t == 0 || outputCount % p.dataFreq == 0) {
        CHECK(sbte_moments(ctx, d_f, d_mom, 1));
        CHECK(sbte_d2h(ctx, mom, d_mom, sizeof(mom)));
        if (isnan(mom[0])) { printf("nan detected\n"); exit(0); }
        fprintf(out, "%le ", p.dt * t);
        if (of.dens) fprintf(out, "%le ", mom[0]);
        if (of.vel) fprintf(out, "%le ", mom[1]);
        if (of.temp) fprintf(out, "%le ", mom[4]);
        if (of.pres) fprintf(out, "%le ", mom[7]);
        if (of.ent) fprintf(out, " %le ", mom[6] / mom[5]);
        if (of.slice) {
          CHECK(sbte_d2h(ctx, f, d_f, sizeof(double) * n3));
          for (l = 0; l < p.N; l++) fprintf(out, " %le ", f[p.N / 2 + p.N * (p.N / 2 + p.N * l)]);
        }
        fprintf(out, "\n");
        outputCount = 0;
      }
    }
    sbte_dev_free(d_f); sbte_dev_free(d_mom); free(f);
  } else {
    /* ---------------- space inhomogeneous: exec/boltz.c:254-395 ----------------
     * The reference runs one MPI rank per block of cells.  Here one process drives SBTE_GPUS GPUs (default 1):
     * "rank" r is GPU r with its own context, weight copy and slab; the blocks may be uneven (the reference needs
     * nX % ranks == 0, src/mesh_setup.c:46-53); ghost cells are never sent -- the stencil kernels read the
     * neighbouring GPU's cells over NVLink (sbte_slab_peer_attach) -- and every step is one graph replay per GPU. */
    int nX, G = 1, r;
    double *x, *dx, *slab_h, *mom;
    sbte_ctx **ctxs;
    sbte_slab **slabs;
    int *lo, *cnt;
    printf("Loading mesh\n");
    make_mesh(p.meshFile, p.order, &nX, &x, &dx);
    if (getenv("SBTE_GPUS")) G = atoi(getenv("SBTE_GPUS"));
    if (G > sbte_device_count()) G = sbte_device_count();
    while (G > 1 && nX / G < 2 * p.order) G--;
    if (G < 1) G = 1;
    ctxs = malloc(sizeof(*ctxs) * G); slabs = malloc(sizeof(*slabs) * G);
    lo = malloc(sizeof(int) * (G + 1)); cnt = malloc(sizeof(int) * G);
    for (r = 0, lo[0] = 0; r < G; r++) { cnt[r] = nX / G + (r < nX % G ? 1 : 0); lo[r + 1] = lo[r] + cnt[r]; }
    ctxs[0] = ctx;
    for (r = 1; r < G; r++) {
      CHECK(sbte_create(&ctxs[r], p.N, p.L_v, v, eta, r));
      setup_weights(ctxs[r], &p, 0);
    }
    if (G > 1) printf("Running on %d GPUs, %d-%d cells each\n", G, cnt[G - 1], cnt[0]);
    slab_h = malloc(sizeof(double) * (size_t)(nX + 2 * p.order) * n3);
    mom = malloc(sizeof(double) * (size_t)nX * 8);
    init_inhom(&p, v, nX, slab_h);
    int t0 = 0;
    double total_start = now_s(), write_start = now_s();
    if (p.restart) {
      printf("Loading from previously generated data\n");   /* the loop resumes AT the stored counter, as exec/boltz.c does */
      for (r = 0; r < G; r++) load_restart(argv[1], slab_h + (size_t)lo[r] * n3, cnt[r], p.order, n3, &t0, r);
    }
    for (r = 0; r < G; r++) {
      /* rank r owns cells lo[r] .. lo[r+1]-1: its slab is that window of the global one plus `order` ghosts per side */
      CHECK(sbte_slab_create(ctxs[r], &slabs[r], cnt[r], p.order, x + lo[r], dx + lo[r], p.initFlag, p.dt, r, G));
      CHECK(sbte_slab_upload(slabs[r], slab_h + (size_t)lo[r] * n3));
    }
    if (G > 1) {
      const int ring = (p.initFlag == 6 && p.order == 1);   /* periodic: src/transportroutines.c:156-172 */
      for (r = 0; r + 1 < G; r++) {
        CHECK(sbte_enable_peer_access(ctxs[r], ctxs[r + 1]));
        CHECK(sbte_slab_peer_attach(slabs[r], 1, slabs[r + 1]));
        CHECK(sbte_slab_peer_attach(slabs[r + 1], 0, slabs[r]));
      }
      if (ring) {
        CHECK(sbte_enable_peer_access(ctxs[0], ctxs[G - 1]));
        CHECK(sbte_slab_peer_attach(slabs[0], 0, slabs[G - 1]));
        CHECK(sbte_slab_peer_attach(slabs[G - 1], 1, slabs[0]));
      }
      for (r = 0; r < G; r++) CHECK(sbte_slab_set_peer_halo(slabs[r], 1));
    }
    if (!p.restart) {
      fprintf(out, "#Time Position ");
      if (of.dens) fprintf(out, "Density ");
      if (of.vel) fprintf(out, "Velocity_x ");
      if (of.temp) fprintf(out, "Temperature ");
      if (of.pres) fprintf(out, "Pressure ");
      fprintf(out, "\n");
    }
    /* exec/boltz.c:255-394 */
#define WRITE_STATE(TIME)                                                                         \
  do {                                                                                            \
    for (r = 0; r < G; r++) CHECK(sbte_slab_moments(slabs[r], mom + (size_t)lo[r] * 8));          \
    for (l = 0; l < nX; l++) {                                                                    \
      const double *m = mom + 8 * l;                                                              \
      if (isnan(m[0])) { printf("nan detected in cell %d \n", l + p.order); exit(0); }            \
      fprintf(out, "%le %le ", (TIME), x[l + p.order]);                                           \
      if (of.dens) fprintf(out, "%le ", m[0]);                                                    \
      if (of.vel) fprintf(out, "%le ", m[1]);                                                     \
      if (of.temp) fprintf(out, "%le ", m[4]);                                                    \
      if (of.pres) fprintf(out, "%le ", m[7]);                                                    \
      fprintf(out, "\n");                                                                         \
    }                                                                                             \
  } while (0)
    if (!p.restart) WRITE_STATE(0.0);
    t = t0;
    while (t < p.nT) {
      /* run up to the next output step; every GPU in its own host thread */
      int chunk = p.dataFreq - outputCount % p.dataFreq, i;
      if (chunk > p.nT - t) chunk = p.nT - t;
      for (i = 0; i < chunk; i++) printf("In step %d of %d\n", t + 1 + i, p.nT);
      run_steps(slabs, ctxs, G, p.Kn, chunk);
      t += chunk;
      outputCount += chunk;
      if (outputCount % p.dataFreq == 0) {
        if (p.restart_time > 0) {
          /* wall-clock triggered checkpoint (:361-388): stop when another output interval would not fit */
          const double write_time = now_s() - write_start, tot_time = now_s() - total_start;
          write_start = now_s();
          if (tot_time + write_time > 0.95 * p.restart_time) {
            printf("RESTART TIME REACHED - STORING CURRENT DISTRIBUTION DATA\n");
            for (r = 0; r < G; r++) {   /* one file per rank, as store_restart writes them under MPI (src/restart.c:24-63) */
              double *tmp = malloc(sizeof(double) * (size_t)(cnt[r] + 2 * p.order) * n3);
              CHECK(sbte_slab_download(slabs[r], tmp));
              store_restart(argv[1], tmp, cnt[r], p.order, n3, t - 1, r);   /* the counter of the last completed step */
              free(tmp);
            }
            fclose(out);
            for (r = 0; r < G; r++) sbte_slab_destroy(slabs[r]);
            for (r = 0; r < G; r++) sbte_destroy(ctxs[r]);
            return 0;
          }
        }
        WRITE_STATE(p.dt * t);
        outputCount = 0;
      }
    }
    for (r = 0; r < G; r++) sbte_slab_destroy(slabs[r]);
    for (r = 1; r < G; r++) sbte_destroy(ctxs[r]);
    free(slab_h); free(mom); free(x); free(dx); free(ctxs); free(slabs); free(lo); free(cnt);

  }
  printf("Wrapping up\n");
  fclose(out);
  sbte_destroy(ctx);
  free(v); free(eta);
  return 0;
}
