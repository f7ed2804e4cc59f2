/* TEST-ONLY driver: executes the .Call entry points of r/src/rshim.c against the fake R runtime
 * (fake_r.c) and the real libsoundgen_b200.so.  Inputs come from a text file written by the Python test,
 * outputs go to another one; the test compares them with the ctypes path and checks what is printed here
 * about PROTECT balance and error handling.   usage: run_shim <cpu|gpu> <inputs> <outputs> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "R.h"

void fake_r_init(void);
int fake_r_protect_depth(void);
const char *fake_r_last_error(void);
int fake_r_try(void (*f)(void *), void *arg);
SEXP fake_r_real(const double *, R_xlen_t);
SEXP fake_r_int(const int *, R_xlen_t);
SEXP fake_r_matrix(const double *, int, int);
SEXP fake_r_named_list(const char **, const double *, int);
SEXP fake_r_list_get(SEXP, const char *);

SEXP sg_generate_harmonics(SEXP, SEXP, SEXP, SEXP);
SEXP sg_generate_noise(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP sg_get_rolloff(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP sg_get_spectral_envelope(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP sg_filter(SEXP, SEXP, SEXP, SEXP);
SEXP sg_soundgen(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP sg_soundgen_batch(SEXP, SEXP, SEXP);
void R_init_soundgen(DllInfo *);

/* ---- inputs: "name n v1 ... vn" records ---- */
typedef struct { char name[64]; int n; double *v; } Rec;
static Rec recs[64];
static int nrec = 0;
static void load(const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) { perror(path); exit(2); }
  while (nrec < 64 && fscanf(f, "%63s %d", recs[nrec].name, &recs[nrec].n) == 2) {
    recs[nrec].v = (double *)calloc((size_t)recs[nrec].n + 1, sizeof(double));
    for (int i = 0; i < recs[nrec].n; i++) if (fscanf(f, "%lf", &recs[nrec].v[i]) != 1) exit(3);
    nrec++;
  }
  fclose(f);
}
static Rec *get(const char *name) {
  for (int i = 0; i < nrec; i++) if (!strcmp(recs[i].name, name)) return &recs[i];
  fprintf(stderr, "missing input %s\n", name); exit(4);
}
static FILE *out;
static void put_real(const char *name, SEXP s) {
  fprintf(out, "%s %ld", name, (long)XLENGTH(s));
  for (R_xlen_t i = 0; i < XLENGTH(s); i++) fprintf(out, " %.17g", REAL(s)[i]);
  fprintf(out, "\n");
}
static void put_int(const char *name, SEXP s) {
  fprintf(out, "%s %ld", name, (long)XLENGTH(s));
  for (R_xlen_t i = 0; i < XLENGTH(s); i++) fprintf(out, " %d", INTEGER(s)[i]);
  fprintf(out, "\n");
}
static SEXP scalar(double v) { return fake_r_real(&v, 1); }
static SEXP iscalar(int v) { return fake_r_int(&v, 1); }

static const char *HP[] = {"nonlinBalance", "jitterDep", "shimmerDep", "samplingRate", "subDep", "subFreq"};
static void t_rolloff(void *u) {
  (void)u;
  Rec *p = get("rolloff_pitch");
  double ro = -12, roct = -2, rk = -6;
  SEXP m = sg_get_rolloff(fake_r_real(p->v, p->n), iscalar(20), scalar(ro), scalar(roct), scalar(rk), scalar(0), scalar(3),
                          R_NilValue, scalar(200), scalar(-120), scalar(16000));
  put_real("rolloff", m);
  fprintf(out, "rolloff_dim 2 %d %d\n", Rf_nrows(m), Rf_ncols(m));
  SEXP dn = Rf_getAttrib(m, R_DimNamesSymbol);
  fprintf(out, "rolloff_lastname 1 %s\n", CHAR(STRING_ELT(VECTOR_ELT(dn, 0), Rf_nrows(m) - 1)));
}
static void t_filter(void *u) {
  (void)u;
  Rec *s = get("filter_sound"), *e = get("filter_env");
  put_real("filter", sg_filter(fake_r_real(s->v, s->n), fake_r_real(e->v, e->n), iscalar(800), scalar(75)));
}
static void t_harm(void *u) {
  (void)u;
  Rec *p = get("harm_pitch"), *z = get("harm_z");
  const double vals[] = {100, 1.0, 5.0, 16000, 60, 80};
  SEXP r = sg_generate_harmonics(fake_r_real(p->v, p->n), fake_r_named_list(HP, vals, 6), fake_r_real(z->v, z->n), R_NilValue);
  put_real("harm_wave", fake_r_list_get(r, "waveform"));
  put_int("harm_z_used", fake_r_list_get(r, "z_used"));
  put_int("harm_gc", fake_r_list_get(r, "gc"));
}
static void t_harm_fail(void *u) {       /* a 3-point contour: the reference's stop('Failed to generate ...') */
  (void)u;
  const double p[3] = {100, 100, 100}, z[4] = {0, 0, 0, 0};
  const double vals[] = {0, 0, 0, 16000, 0, 100};
  sg_generate_harmonics(fake_r_real(p, 3), fake_r_named_list(HP, vals, 6), fake_r_real(z, 4), R_NilValue);
}
static const char *NP[] = {"rolloffNoise", "attackLen", "windowLength_points", "samplingRate", "overlap", "throwaway"};
static void t_noise(void *u) {
  (void)u;
  Rec *uu = get("noise_u");
  const double an[4] = {0, 300, -20, -10};          /* column-major (time, value) */
  const double vals[] = {-6, 10, 800, 16000, 75, -120};
  put_real("noise", sg_generate_noise(iscalar(3000), fake_r_matrix(an, 2, 2), R_NilValue, fake_r_real(uu->v, uu->n),
                                      R_NilValue, fake_r_named_list(NP, vals, 6)));
}
static void t_noise_badfilter(void *u) {  /* Rf_error before any native resource exists */
  (void)u;
  const double an[4] = {0, 300, -20, -10}, flt[10] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, uu[8] = {.5, .5, .5, .5, .5, .5, .5, .5};
  const double vals[] = {-6, 10, 800, 16000, 75, -120};
  sg_generate_noise(iscalar(3000), fake_r_matrix(an, 2, 2), R_NilValue, fake_r_real(uu, 8), fake_r_matrix(flt, 10, 1),
                    fake_r_named_list(NP, vals, 6));
}
static void t_env(void *u) {
  (void)u;
  const double fm[12] = {0, 0, 0, 860, 1280, 2900, 30, 40, 25, 120, 120, 200};   /* 3 x 4 column-major */
  const int fn[3] = {1, 1, 1};
  const char *nm[] = {"vocalTract", "samplingRate"};
  const double vals[] = {15.5, 16000};
  put_real("env", sg_get_spectral_envelope(iscalar(400), iscalar(5), fake_r_matrix(fm, 3, 4), fake_r_int(fn, 3), iscalar(0),
                                           R_NilValue, fake_r_named_list(nm, vals, 2)));
}
static void t_filter_short(void *u) {
  (void)u;
  const double s[4] = {1, 2, 3, 4}, e[2] = {1, 1};
  sg_filter(fake_r_real(s, 4), fake_r_real(e, 2), iscalar(4), scalar(75));
}
static void t_soundgen(void *u) {
  (void)u;
  Rec *sd = get("seed");
  int *seed = (int *)calloc(626, sizeof(int));
  for (int i = 0; i < 626; i++) seed[i] = (int)sd->v[i];
  const char *nm[] = {"sylLen", "sylLenDep"};
  const double vals[] = {300, .02};
  const double pa[8] = {0, .1, .9, 1, 100, 150, 135, 100}, na[4] = {0, 300, -120, -120}, ma[4] = {0, 1, .5, .5};
  SEXP anchors = Rf_allocVector(VECSXP, 6);
  SET_VECTOR_ELT(anchors, 0, fake_r_matrix(pa, 4, 2));
  SET_VECTOR_ELT(anchors, 2, fake_r_matrix(na, 2, 2));
  SET_VECTOR_ELT(anchors, 3, fake_r_matrix(ma, 2, 2));
  const double f1[4] = {0, 860, 30, 120}, f2[4] = {0, 1280, 40, 120}, f3[4] = {0, 2900, 25, 200};
  SEXP fl = Rf_allocVector(VECSXP, 3);
  SET_VECTOR_ELT(fl, 0, fake_r_matrix(f1, 1, 4)); SET_VECTOR_ELT(fl, 1, fake_r_matrix(f2, 1, 4)); SET_VECTOR_ELT(fl, 2, fake_r_matrix(f3, 1, 4));
  SEXP r = sg_soundgen(fake_r_named_list(nm, vals, 2), anchors, fl, R_NilValue, fake_r_named_list(nm, vals, 0), fake_r_int(seed, 626));
  put_real("soundgen_wave", fake_r_list_get(r, "waveform"));
  put_int("soundgen_seed", fake_r_list_get(r, "seed"));
  put_int("soundgen_status", fake_r_list_get(r, "status"));
}

static void t_soundgen_batch(void *u) {   /* set.seed(1) and set.seed(2) versions of soundgen(sylLen = 300) as one batch */
  (void)u;
  const char *nm[] = {"sylLen", "sylLenDep"};
  const double vals[] = {300, .02};
  const double pa[8] = {0, .1, .9, 1, 100, 150, 135, 100}, na[4] = {0, 300, -120, -120}, ma[4] = {0, 1, .5, .5};
  const double f1[4] = {0, 860, 30, 120}, f2[4] = {0, 1280, 40, 120}, f3[4] = {0, 2900, 25, 200};
  SEXP calls = Rf_allocVector(VECSXP, 2);
  for (int i = 0; i < 2; i++) {
    SEXP anchors = Rf_allocVector(VECSXP, 6);
    SET_VECTOR_ELT(anchors, 0, fake_r_matrix(pa, 4, 2));
    SET_VECTOR_ELT(anchors, 2, fake_r_matrix(na, 2, 2));
    SET_VECTOR_ELT(anchors, 3, fake_r_matrix(ma, 2, 2));
    SEXP fl = Rf_allocVector(VECSXP, 3);
    SET_VECTOR_ELT(fl, 0, fake_r_matrix(f1, 1, 4)); SET_VECTOR_ELT(fl, 1, fake_r_matrix(f2, 1, 4)); SET_VECTOR_ELT(fl, 2, fake_r_matrix(f3, 1, 4));
    SEXP c = Rf_allocVector(VECSXP, 4);
    SET_VECTOR_ELT(c, 0, fake_r_named_list(nm, vals, 2)); SET_VECTOR_ELT(c, 1, anchors); SET_VECTOR_ELT(c, 2, fl);
    SET_VECTOR_ELT(calls, i, c);
  }
  const int seeds[2] = {1, 2};
  SEXP r = sg_soundgen_batch(calls, fake_r_named_list(nm, vals, 0), fake_r_int(seeds, 2));
  put_real("batch_wave_1", VECTOR_ELT(r, 0));
  put_real("batch_wave_2", VECTOR_ELT(r, 1));
}

static int run(const char *name, void (*f)(void *), int expect_error) {
  int failed = fake_r_try(f, NULL);
  printf("%s: %s%s%s; protect depth %d\n", name, failed ? "Rf_error: " : "ok", failed ? fake_r_last_error() : "",
         (failed != expect_error) ? "  <-- UNEXPECTED" : "", fake_r_protect_depth());
  return (failed != expect_error) || fake_r_protect_depth() != 0;
}

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: run_shim <cpu|gpu> <inputs> <outputs>\n"); return 2; }
  const int gpu = !strcmp(argv[1], "gpu");
  fake_r_init();
  load(argv[2]);
  out = fopen(argv[3], "w");
  R_init_soundgen(NULL);
  int bad = 0;
  bad |= run("noise_badfilter", t_noise_badfilter, 1);
  bad |= run("filter_short", t_filter_short, 1);
  if (!gpu) {            /* without a device every compute entry must fail cleanly: no CPU fallback */
    bad |= run("rolloff_nodevice", t_rolloff, 1);
    bad |= run("harmonics_nodevice", t_harm, 1);
    bad |= run("soundgen_nodevice", t_soundgen, 1);
  } else {
    bad |= run("rolloff", t_rolloff, 0);
    bad |= run("filter", t_filter, 0);
    bad |= run("harmonics", t_harm, 0);
    bad |= run("harmonics_fail", t_harm_fail, 1);
    bad |= run("noise", t_noise, 0);
    bad |= run("envelope", t_env, 0);
    bad |= run("soundgen", t_soundgen, 0);
    bad |= run("soundgen_batch", t_soundgen_batch, 0);
  }
  fclose(out);
  printf("%s\n", bad ? "SHIM TEST FAILED" : "SHIM TEST PASSED");
  return bad;
}
