/*
 * rshim.c -- .Call bindings of libsoundgen_b200 for the soundgen R package.
 *
 * Marshalling only: no arithmetic on samples happens here (BASELINE north_star: "no Rcpp-side
 * math").  Every entry point converts R vectors to the POD structs of include/soundgen_b200.h,
 * calls the C ABI and wraps the result in a fresh R vector.  Errors of the library become R
 * errors AFTER all native resources are released (Rf_error long-jumps).
 *
 * Replaces (reference file:line):
 *   sg_generate_harmonics   <- generateHarmonics        R/source.R:173-471
 *   sg_generate_noise       <- generateNoise            R/source.R:57-138
 *   sg_get_rolloff          <- getRolloff               R/sourceSpectrum.R:71-186
 *   sg_get_spectral_envelope<- getSpectralEnvelope      R/sourceSpectrum.R:417-566 (deterministic part)
 *   sg_filter               <- filter block of soundgen R/soundgen.R:743-807 (+ seewave stft/istft)
 *   sg_soundgen             <- soundgen() itself        R/soundgen.R:279-849 (host stage in the library's
 *                              front-end, drawing from R's own stream: .Random.seed goes in and comes back)
 *
 * Build: see src/Makevars (links against libsoundgen_b200.so built by __graft_entry__.build()).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include <math.h>

#include "soundgen_b200.h"

static double num(SEXP list, const char *name, double dflt) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(list); i++)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return Rf_asReal(VECTOR_ELT(list, i));
  return dflt;
}

static void fail_if(int rc, sgb_batch *b) {
  if (rc >= 0) return;
  char msg[512];
  strncpy(msg, sgb_last_error(), sizeof msg - 1);
  msg[sizeof msg - 1] = 0;
  if (b) sgb_batch_destroy(b);
  if (rc == SGB_ERR_SYNTH) Rf_error("Failed to generate the new syllable!");   /* soundgen.R:624-626 */
  Rf_error("soundgen_b200: %s", msg);
}

/* pars: named list of the numeric arguments of generateHarmonics (source.R:173-205);
 * pitch: numeric contour; z: standard normals drawn by the caller (rnorm(cap));
 * ampl: NULL or a 2-column numeric matrix (time, value).
 * Returns list(waveform = numeric, z_used = integer, gc = integer, gc_upsampled = integer). */
SEXP sg_generate_harmonics(SEXP pitch, SEXP pars, SEXP z, SEXP ampl) {
  sgb_syllable s;
  memset(&s, 0, sizeof s);
  s.kind = 1;
  s.pitch_len = (int32_t)XLENGTH(pitch);
  s.z_cap = (int32_t)XLENGTH(z);
#define P(f, d) s.f = num(pars, #f, d)
  P(attackLen, 50); P(nonlinBalance, 0); P(jitterDep, 0); P(jitterLen, 1); P(vibratoFreq, 100);
  P(vibratoDep, 0); P(shimmerDep, 0); P(rolloff, -18); P(rolloffOct, -2); P(rolloffKHz, -6);
  P(rolloffParab, 0); P(rolloffParabHarm, 3); P(rolloff_perAmpl, 12); P(temperature, 0);
  P(pitchDriftDep, .5); P(pitchDriftFreq, .125); P(randomWalk_trendStrength, .5); P(shortestEpoch, 300);
  P(subFreq, 100); P(subDep, 0); P(samplingRate, 16000); P(pitchFloor, 75); P(pitchCeiling, 3500);
  P(pitchSamplingRate, 3500); P(throwaway, -120);
#undef P
  double *anch = NULL;
  int n_anch = 0;
  if (!Rf_isNull(ampl)) {   /* column-major (time, value) matrix -> interleaved pairs */
    n_anch = Rf_nrows(ampl);
    anch = (double *)R_alloc(2 * (size_t)n_anch, sizeof(double));
    for (int i = 0; i < n_anch; i++) { anch[2 * i] = REAL(ampl)[i]; anch[2 * i + 1] = REAL(ampl)[n_anch + i]; }
    s.ampl_n = n_anch;
  }
  sgb_envelope env;
  memset(&env, 0, sizeof env);
  env.formantDep = 1; env.vocalTract = R_NaN; env.samplingRate = s.samplingRate; env.speedSound = 35400;
  env.smoothLinearFactor = 1;
  sgb_bout bout;
  memset(&bout, 0, sizeof bout);
  bout.syl_end = 1; bout.wl = (int32_t)(floor(50.0 / 1000 * s.samplingRate / 2) * 2);
  bout.overlap = 75; bout.samplingRate = s.samplingRate; bout.throwaway = s.throwaway;
  sgb_call call = {0, 1};
  sgb_batch_desc d;
  memset(&d, 0, sizeof d);
  d.n_calls = d.n_bouts = d.n_syllables = d.n_envelopes = 1;
  d.calls = &call; d.bouts = &bout; d.syllables = &s; d.envelopes = &env;
  d.pitch = REAL(pitch); d.n_pitch = XLENGTH(pitch);
  d.anchors = anch; d.n_anchors = n_anch;
  d.z = REAL(z); d.n_z = XLENGTH(z);

  sgb_batch *b = NULL;
  fail_if(sgb_batch_create(&b), NULL);
  fail_if(sgb_batch_upload(b, &d), b);
  fail_if(sgb_batch_run(b, NULL), b);
  sgb_syl_artefacts art;
  fail_if(sgb_batch_artefacts(b, 0, &art), b);
  fail_if(art.status, b);
  int64_t len = 0;
  fail_if(sgb_batch_syllable_len(b, 0, &len), b);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SEXP w = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)len));
  SEXP gc = PROTECT(Rf_allocVector(INTSXP, art.nGC));
  SEXP gcu = PROTECT(Rf_allocVector(INTSXP, art.nGC + 1));
  int rc = sgb_batch_syllable_fetch(b, 0, REAL(w), len);
  if (rc >= 0) rc = sgb_batch_artefact_ints(b, 0, 0, INTEGER(gc), art.nGC);
  if (rc >= 0) rc = sgb_batch_artefact_ints(b, 0, 1, INTEGER(gcu), art.nGC + 1);
  if (rc < 0) { UNPROTECT(4); fail_if(rc, b); }
  sgb_batch_destroy(b);
  SET_VECTOR_ELT(out, 0, w);
  SET_VECTOR_ELT(out, 1, Rf_ScalarInteger(art.z_used));
  SET_VECTOR_ELT(out, 2, gc);
  SET_VECTOR_ELT(out, 3, gcu);
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 4));
  SET_STRING_ELT(nm, 0, Rf_mkChar("waveform")); SET_STRING_ELT(nm, 1, Rf_mkChar("z_used"));
  SET_STRING_ELT(nm, 2, Rf_mkChar("gc")); SET_STRING_ELT(nm, 3, Rf_mkChar("gc_upsampled"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(5);
  return out;
}

/* generateNoise (source.R:57-138).  len: integer; anchors: 2-column matrix (time ms, value dB) or NULL
 * when `strength` (numeric[len], the contour evaluated in R, e.g. by loess) is given; u: runif(nr * nc)
 * drawn by the caller; filter: NULL or the nr x k matrix filterNoise; pars: named list of
 * rolloffNoise, attackLen, windowLength_points, samplingRate, overlap, throwaway.  Returns numeric[len]. */
SEXP sg_generate_noise(SEXP len, SEXP anchors, SEXP strength, SEXP u, SEXP filter, SEXP pars) {
  const int L = Rf_asInteger(len);
  const int wl = (int)num(pars, "windowLength_points", 1024);
  sgb_syllable syl;
  memset(&syl, 0, sizeof syl);
  syl.kind = 0; syl.silent_len = 8;                       /* a bout needs a syllable; this one is silence */
  sgb_envelope env[2];
  memset(env, 0, sizeof env);
  for (int i = 0; i < 2; i++) {
    env[i].formantDep = 1; env[i].vocalTract = R_NaN; env[i].samplingRate = num(pars, "samplingRate", 16000);
    env[i].speedSound = 35400; env[i].smoothLinearFactor = 1;
  }
  sgb_noise nz;
  memset(&nz, 0, sizeof nz);
  nz.len = L; nz.insertion = 1; nz.mix = 0; nz.wl = wl; nz.env_id = -1; nz.strength_pre_off = -1;
  nz.rolloffNoise = num(pars, "rolloffNoise", -6); nz.attackLen = num(pars, "attackLen", 10);
  nz.samplingRate = num(pars, "samplingRate", 16000); nz.overlap = num(pars, "overlap", 75);
  /* pools: anchors as interleaved pairs; `pre` = [strength | filter matrix] */
  double *anch = NULL;
  int n_anch = 0;
  if (!Rf_isNull(anchors)) {
    n_anch = Rf_nrows(anchors);
    anch = (double *)R_alloc(2 * (size_t)n_anch, sizeof(double));
    for (int i = 0; i < n_anch; i++) { anch[2 * i] = REAL(anchors)[i]; anch[2 * i + 1] = REAL(anchors)[n_anch + i]; }
    nz.anchor_n = n_anch;
  }
  int64_t n_pre = 0;
  const int64_t n_str = Rf_isNull(strength) ? 0 : (int64_t)XLENGTH(strength);
  const int64_t n_flt = Rf_isNull(filter) ? 0 : (int64_t)XLENGTH(filter);
  double *pre = NULL;
  if (n_str + n_flt > 0) {
    pre = (double *)R_alloc((size_t)(n_str + n_flt), sizeof(double));
    if (n_str) { memcpy(pre, REAL(strength), sizeof(double) * (size_t)n_str); nz.strength_pre_off = 0; nz.anchor_n = 0; }
    if (n_flt) {
      if (Rf_nrows(filter) != wl / 2) Rf_error("filterNoise must have windowLength_points / 2 rows");
      memcpy(pre + n_str, REAL(filter), sizeof(double) * (size_t)n_flt);   /* already column-major */
      env[1].tracks_given = 2; env[1].formant_off = n_str; env[1].nc_fixed = Rf_isMatrix(filter) ? Rf_ncols(filter) : 1;
      nz.env_id = 1;
    }
    n_pre = n_str + n_flt;
  }
  sgb_bout bout;
  memset(&bout, 0, sizeof bout);
  bout.syl_end = 1; bout.noise_end = 1; bout.wl = wl < 4 ? 4 : wl; bout.overlap = 75;
  bout.samplingRate = nz.samplingRate; bout.throwaway = num(pars, "throwaway", -120);
  sgb_call call = {0, 1};
  sgb_batch_desc d;
  memset(&d, 0, sizeof d);
  d.n_calls = d.n_bouts = d.n_syllables = d.n_noises = 1; d.n_envelopes = 2;
  d.calls = &call; d.bouts = &bout; d.syllables = &syl; d.noises = &nz; d.envelopes = env;
  d.anchors = anch; d.n_anchors = n_anch;
  d.u = REAL(u); d.n_u = XLENGTH(u); d.u_is_float = 0;
  d.pre = pre; d.n_pre = n_pre;
  sgb_batch *b = NULL;
  fail_if(sgb_batch_create(&b), NULL);
  fail_if(sgb_batch_upload(b, &d), b);
  fail_if(sgb_batch_run(b, NULL), b);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)L));
  int rc = sgb_batch_noise_fetch(b, 0, REAL(out), L);
  UNPROTECT(1);
  fail_if(rc, b);
  sgb_batch_destroy(b);
  return out;
}

/* getRolloff: returns a matrix rows x nGC with rownames 1..rows (used as times_f0, source.R:401) */
SEXP sg_get_rolloff(SEXP pitch_per_gc, SEXP nHarmonics, SEXP rolloff, SEXP rolloffOct, SEXP rolloffKHz,
                    SEXP rolloffParab, SEXP rolloffParabHarm, SEXP rolloffParabCeiling, SEXP baseline,
                    SEXP throwaway, SEXP samplingRate) {
  int G = (int)XLENGTH(pitch_per_gc), nH = Rf_asInteger(nHarmonics);
  double *tmp = (double *)R_alloc((size_t)G * nH, sizeof(double));
  int32_t rows = 0;
  double ceil_ = Rf_isNull(rolloffParabCeiling) ? -1.0 : Rf_asReal(rolloffParabCeiling);
  fail_if(sgb_get_rolloff(REAL(pitch_per_gc), G, nH, REAL(rolloff), (int)XLENGTH(rolloff), REAL(rolloffOct),
                          (int)XLENGTH(rolloffOct), REAL(rolloffKHz), (int)XLENGTH(rolloffKHz),
                          Rf_asReal(rolloffParab), Rf_asReal(rolloffParabHarm), ceil_, Rf_asReal(baseline),
                          Rf_asReal(throwaway), Rf_asReal(samplingRate), tmp, &rows), NULL);
  SEXP m = PROTECT(Rf_allocMatrix(REALSXP, rows, G));
  for (int g = 0; g < G; g++) memcpy(REAL(m) + (size_t)g * rows, tmp + (size_t)g * nH, sizeof(double) * rows);
  SEXP rn = PROTECT(Rf_allocVector(STRSXP, rows));
  char buf[16];
  for (int i = 0; i < rows; i++) { snprintf(buf, sizeof buf, "%d", i + 1); SET_STRING_ELT(rn, i, Rf_mkChar(buf)); }
  SEXP dn = PROTECT(Rf_allocVector(VECSXP, 2));
  SET_VECTOR_ELT(dn, 0, rn); SET_VECTOR_ELT(dn, 1, R_NilValue);
  Rf_setAttrib(m, R_DimNamesSymbol, dn);
  UNPROTECT(3);
  return m;
}

/* getSpectralEnvelope, deterministic part.  formants: numeric matrix with 4 columns
 * (time, freq, amp, width), the formants stacked row-wise; formant_n: rows per formant;
 * mouth: NULL or 2-column matrix; pars: named list of scalars. */
SEXP sg_get_spectral_envelope(SEXP nr, SEXP nc, SEXP formants, SEXP formant_n, SEXP tracks_given, SEXP mouth,
                              SEXP pars) {
  sgb_envelope e;
  memset(&e, 0, sizeof e);
  e.n_formants = Rf_isNull(formants) ? 0 : (int)XLENGTH(formant_n);
  e.tracks_given = Rf_asLogical(tracks_given) ? 1 : 0;
  e.formantDep = num(pars, "formantDep", 1); e.rolloffLip = num(pars, "rolloffLip", 6);
  e.mouthOpenThres = num(pars, "mouthOpenThres", 0); e.openMouthBoost = num(pars, "openMouthBoost", 0);
  e.vocalTract = num(pars, "vocalTract", R_NaN); e.samplingRate = num(pars, "samplingRate", 16000);
  e.speedSound = num(pars, "speedSound", 35400); e.smoothLinearFactor = num(pars, "smoothLinearFactor", 1);
  double *rows = NULL, *ma = NULL;
  if (e.n_formants > 0) {   /* column-major R matrix -> row-major (time, freq, amp, width) */
    int n = Rf_nrows(formants);
    rows = (double *)R_alloc(4 * (size_t)n, sizeof(double));
    for (int i = 0; i < n; i++) for (int j = 0; j < 4; j++) rows[4 * i + j] = REAL(formants)[(size_t)j * n + i];
  }
  if (!Rf_isNull(mouth)) {
    int n = Rf_nrows(mouth);
    ma = (double *)R_alloc(2 * (size_t)n, sizeof(double));
    for (int i = 0; i < n; i++) { ma[2 * i] = REAL(mouth)[i]; ma[2 * i + 1] = REAL(mouth)[n + i]; }
    e.mouth_n = n;
  }
  int NR = Rf_asInteger(nr), NC = Rf_asInteger(nc);
  SEXP m = PROTECT(Rf_allocMatrix(REALSXP, NR, NC));
  int rc = sgb_get_spectral_envelope(NR, NC, &e, rows, e.n_formants ? INTEGER(formant_n) : NULL, ma, REAL(m));
  UNPROTECT(1);
  fail_if(rc, NULL);
  return m;
}

/* soundFiltered = istft(stft(sound) * envelope) / max  (soundgen.R:743-807) */
SEXP sg_filter(SEXP sound, SEXP envelope, SEXP wl, SEXP overlap) {
  int64_t len = XLENGTH(sound);
  int nInt = Rf_isMatrix(envelope) ? Rf_ncols(envelope) : 1;
  int64_t n = sgb_filter_len(len, Rf_asInteger(wl), Rf_asReal(overlap));
  if (n < 0) Rf_error("soundgen_b200: sound too short to filter");
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
  int rc = sgb_filter(REAL(sound), len, REAL(envelope), nInt, Rf_asInteger(wl), Rf_asReal(overlap), REAL(out), n);
  UNPROTECT(1);
  fail_if(rc, NULL);
  return out;
}

/* ---- soundgen() as one call ------------------------------------------------------------------------
 * args:     named list of the numeric arguments of soundgen() (R/soundgen.R:208-277) and of tempEffects
 *           (sylLenDep, formDrift, ...); anything missing takes the reference's default.
 * anchors:  list of 6: pitchAnchors, pitchAnchorsGlobal, noiseAnchors, mouthAnchors, amplAnchors,
 *           amplAnchorsGlobal -- each NULL (NA) or a numeric matrix with the columns (time, value).
 * formants, formantsNoise: NULL (NA) or a list of numeric matrices with the columns (time, freq, amp, width).
 * opts:     named list: invalidArgAction (0 adjust, 1 abort, 2 ignore), contour_method (0 loess, 1 spline),
 *           sample_rejection (1 under R >= 3.6 with the default sample.kind).
 * seed:     .Random.seed (integer, 626: kind code, mti, mt[624]) of a Mersenne-Twister / Inversion session.
 * Returns list(waveform = numeric, seed = integer(626), status = integer, warnings = character). */
static void anchor_arg(SEXP m, sgb_anchor_arg *a) {
  memset(a, 0, sizeof *a);
  if (Rf_isNull(m)) return;
  int n = Rf_nrows(m);
  a->time = REAL(m); a->value = REAL(m) + n; a->n = n;       /* column-major: the time column, then the values */
}
static sgb_formant_arg *formant_args(SEXP fl, int *n_out) {
  *n_out = 0;
  if (Rf_isNull(fl)) return NULL;
  int nf = (int)XLENGTH(fl);
  sgb_formant_arg *f = (sgb_formant_arg *)R_alloc((size_t)nf, sizeof(sgb_formant_arg));
  for (int i = 0; i < nf; i++) {
    SEXP m = VECTOR_ELT(fl, i);
    int n = Rf_nrows(m);
    f[i].time = REAL(m); f[i].freq = REAL(m) + n; f[i].amp = REAL(m) + 2 * (size_t)n; f[i].width = REAL(m) + 3 * (size_t)n;
    f[i].n_time = f[i].n_freq = f[i].n_amp = f[i].n_width = n;
  }
  *n_out = nf;
  return f;
}

static void fill_args(SEXP args, SEXP anchors, SEXP formants, SEXP formantsNoise, SEXP opts, sgb_soundgen_args *out) {
  sgb_soundgen_args a;
  memset(&a, 0, sizeof a);
#define A(f, d) a.f = num(args, #f, d)
  A(repeatBout, 1); A(nSyl, 1); A(sylLen, 300); A(pauseLen, 200); A(temperature, 0.025); A(maleFemale, 0);
  A(creakyBreathy, 0); A(nonlinBalance, 0); A(nonlinDep, 50); A(jitterLen, 1); A(jitterDep, 3); A(vibratoFreq, 5);
  A(vibratoDep, 0); A(shimmerDep, 0); A(attackLen, 50); A(rolloff, -12); A(rolloffOct, -12); A(rolloffKHz, -6);
  A(rolloffParab, 0); A(rolloffParabHarm, 3); A(rolloffLip, 6); A(formantDep, 1); A(formantDepStoch, 30);
  A(vocalTract, 15.5); A(subFreq, 100); A(subDep, 100); A(shortestEpoch, 300); A(amDep, 0); A(amFreq, 30);
  A(amShape, 0); A(rolloffNoise, -14); A(samplingRate, 16000); A(windowLength, 50); A(overlap, 75);
  A(addSilence, 100); A(pitchFloor, 50); A(pitchCeiling, 3500); A(pitchSamplingRate, 3500); A(throwaway, -120);
#undef A
  static const char *te[8] = {"sylLenDep", "formDrift", "formDisp", "pitchDriftDep", "pitchDriftFreq",
                              "pitchAnchorsDep", "noiseAnchorsDep", "amplAnchorsDep"};
  for (int i = 0; i < 8; i++) a.tempEffects[i] = num(args, te[i], R_NaN);
  if (XLENGTH(anchors) != 6) Rf_error("soundgen_b200: `anchors` must be a list of 6");
  sgb_anchor_arg *slots[6] = {&a.pitchAnchors, &a.pitchAnchorsGlobal, &a.noiseAnchors, &a.mouthAnchors,
                              &a.amplAnchors, &a.amplAnchorsGlobal};
  for (int i = 0; i < 6; i++) anchor_arg(VECTOR_ELT(anchors, i), slots[i]);
  a.formants = formant_args(formants, &a.n_formants);
  a.formantsNoise = formant_args(formantsNoise, &a.n_formantsNoise);
  a.invalidArgAction = (int)num(opts, "invalidArgAction", 0);
  a.contour_method = (int)num(opts, "contour_method", 0);
  a.sample_rejection = (int)num(opts, "sample_rejection", 0);
  a.device_pitch = 1;
  *out = a;
}

SEXP sg_soundgen(SEXP args, SEXP anchors, SEXP formants, SEXP formantsNoise, SEXP opts, SEXP seed) {
  sgb_soundgen_args a;
  fill_args(args, anchors, formants, formantsNoise, opts, &a);
  if (XLENGTH(seed) != 626) Rf_error("soundgen_b200: .Random.seed of a Mersenne-Twister session (626 integers) expected");
  a.rng_mode = 1;
  a.rng_state = INTEGER(seed) + 1;

  sgb_frontend *fe = NULL;
  sgb_batch *b = NULL;
  int rc = sgb_frontend_create(&fe, 0);
  if (rc >= 0) rc = sgb_frontend_add(fe, &a);
  if (rc >= 0) rc = sgb_batch_create(&b);
  double *wave = NULL;
  int64_t total = 0;
  while (rc >= 0) {                       /* one round per bout whose stochastic filter needs the device's frame count */
    sgb_batch_desc d;
    int32_t nsub = 0;
    rc = sgb_frontend_round_begin(fe, &d, &nsub);
    if (rc < 0 || nsub == 0) break;
    if ((rc = sgb_batch_upload(b, &d)) < 0) break;
    if ((rc = sgb_batch_run_begin(b)) < 0) break;
    if ((rc = sgb_frontend_resolve(fe, b)) < 0) break;
    if ((rc = sgb_batch_run_finish(b, NULL)) < 0) break;
    if ((rc = sgb_frontend_round_end(fe, b)) < 0) break;
    int64_t len = 0;
    if ((rc = sgb_batch_lengths(b, &len)) < 0) break;
    double *grown = (double *)R_alloc((size_t)(total + len + 1), sizeof(double));
    if (total) memcpy(grown, wave, sizeof(double) * (size_t)total);
    wave = grown;
    if (len > 0 && (rc = sgb_batch_fetch_f64(b, wave + total, len)) < 0) break;
    total += len;
  }
  int32_t status = 0;
  int32_t state[625];
  char warn[1024];
  warn[0] = 0;
  if (rc >= 0) {
    sgb_frontend_status(fe, &status);
    sgb_frontend_rng_state(fe, 0, state);
    strncpy(warn, sgb_frontend_warnings(fe, 0), sizeof warn - 1);
    warn[sizeof warn - 1] = 0;
  }
  char msg[512];
  msg[0] = 0;
  if (rc < 0) { strncpy(msg, sgb_last_error(), sizeof msg - 1); msg[sizeof msg - 1] = 0; }
  if (b) sgb_batch_destroy(b);
  if (fe) sgb_frontend_destroy(fe);
  if (rc < 0) Rf_error("soundgen_b200: %s", msg);
  if (status == SGB_ERR_SYNTH) Rf_error("Failed to generate the new syllable!");   /* soundgen.R:624-626 */
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SEXP w = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)total));
  if (total) memcpy(REAL(w), wave, sizeof(double) * (size_t)total);
  SEXP sd = PROTECT(Rf_allocVector(INTSXP, 626));
  INTEGER(sd)[0] = INTEGER(seed)[0];
  memcpy(INTEGER(sd) + 1, state, sizeof state);
  SEXP wn = PROTECT(Rf_allocVector(STRSXP, 1));
  SET_STRING_ELT(wn, 0, Rf_mkChar(warn));
  SET_VECTOR_ELT(out, 0, w);
  SET_VECTOR_ELT(out, 1, sd);
  SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(status));
  SET_VECTOR_ELT(out, 3, wn);
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 4));
  SET_STRING_ELT(nm, 0, Rf_mkChar("waveform")); SET_STRING_ELT(nm, 1, Rf_mkChar("seed"));
  SET_STRING_ELT(nm, 2, Rf_mkChar("status")); SET_STRING_ELT(nm, 3, Rf_mkChar("warnings"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(5);
  return out;
}

/* soundgen_batch(): calls = list of list(args, anchors, formants, formantsNoise) as for sg_soundgen, seeds =
 * integer vector (call i runs as after set.seed(seeds[i])).  All calls go through the device as ONE batch
 * (plus one more round per bout whose stochastic filter waits for the device's frame count).
 * Returns a list of numeric waveforms; a call the reference would stop() on comes back as NULL. */
SEXP sg_soundgen_batch(SEXP calls, SEXP opts, SEXP seeds) {
  const int n = (int)XLENGTH(calls);
  if (XLENGTH(seeds) != n) Rf_error("soundgen_b200: one seed per call expected");
  sgb_frontend *fe = NULL;
  sgb_batch *b = NULL;
  int rc = sgb_frontend_create(&fe, 0);
  for (int i = 0; i < n && rc >= 0; i++) {
    SEXP c = VECTOR_ELT(calls, i);
    sgb_soundgen_args a;
    fill_args(VECTOR_ELT(c, 0), VECTOR_ELT(c, 1), VECTOR_ELT(c, 2), VECTOR_ELT(c, 3), opts, &a);
    a.rng_mode = 0;
    a.seed = (uint32_t)INTEGER(seeds)[i];
    rc = sgb_frontend_add(fe, &a);
  }
  if (rc >= 0) rc = sgb_batch_create(&b);
  double **part = (double **)R_alloc((size_t)n + 1, sizeof(double *));
  int64_t *plen = (int64_t *)R_alloc((size_t)n + 1, sizeof(int64_t));
  for (int i = 0; i < n; i++) { part[i] = NULL; plen[i] = 0; }
  while (rc >= 0) {
    sgb_batch_desc d;
    int32_t nsub = 0;
    rc = sgb_frontend_round_begin(fe, &d, &nsub);
    if (rc < 0 || nsub == 0) break;
    if ((rc = sgb_batch_upload(b, &d)) < 0) break;
    if ((rc = sgb_batch_run_begin(b)) < 0) break;
    if ((rc = sgb_frontend_resolve(fe, b)) < 0) break;
    if ((rc = sgb_batch_run_finish(b, NULL)) < 0) break;
    if ((rc = sgb_frontend_round_end(fe, b)) < 0) break;
    int64_t *lens = (int64_t *)R_alloc((size_t)nsub, sizeof(int64_t));
    int32_t *owner = (int32_t *)R_alloc((size_t)nsub, sizeof(int32_t));
    if ((rc = sgb_batch_lengths(b, lens)) < 0) break;
    if ((rc = sgb_frontend_round_calls(fe, owner)) < 0) break;
    int64_t total = 0;
    for (int k = 0; k < nsub; k++) total += lens[k];
    double *all = (double *)R_alloc((size_t)total + 1, sizeof(double));
    if (total > 0 && (rc = sgb_batch_fetch_f64(b, all, total)) < 0) break;
    int64_t off = 0;
    for (int k = 0; k < nsub; k++) {       /* append this round's piece to its call */
      const int c = owner[k];
      double *grown = (double *)R_alloc((size_t)(plen[c] + lens[k]) + 1, sizeof(double));
      if (plen[c]) memcpy(grown, part[c], sizeof(double) * (size_t)plen[c]);
      memcpy(grown + plen[c], all + off, sizeof(double) * (size_t)lens[k]);
      part[c] = grown; plen[c] += lens[k]; off += lens[k];
    }
  }
  int32_t *status = (int32_t *)R_alloc((size_t)n + 1, sizeof(int32_t));
  if (rc >= 0) rc = sgb_frontend_status(fe, status);
  char msg[512];
  msg[0] = 0;
  if (rc < 0) { strncpy(msg, sgb_last_error(), sizeof msg - 1); msg[sizeof msg - 1] = 0; }
  if (b) sgb_batch_destroy(b);
  if (fe) sgb_frontend_destroy(fe);
  if (rc < 0) Rf_error("soundgen_b200: %s", msg);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n));
  for (int i = 0; i < n; i++) {
    if (status[i] != 0) continue;                       /* stays NULL */
    SEXP w = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)plen[i]));
    if (plen[i]) memcpy(REAL(w), part[i], sizeof(double) * (size_t)plen[i]);
    SET_VECTOR_ELT(out, i, w);
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef call_methods[] = {
    {"sg_generate_harmonics", (DL_FUNC)&sg_generate_harmonics, 4},
    {"sg_generate_noise", (DL_FUNC)&sg_generate_noise, 6},
    {"sg_get_rolloff", (DL_FUNC)&sg_get_rolloff, 11},
    {"sg_get_spectral_envelope", (DL_FUNC)&sg_get_spectral_envelope, 7},
    {"sg_filter", (DL_FUNC)&sg_filter, 4},
    {"sg_soundgen", (DL_FUNC)&sg_soundgen, 6},
    {"sg_soundgen_batch", (DL_FUNC)&sg_soundgen_batch, 3},
    {NULL, NULL, 0}};

void R_init_soundgen(DllInfo *dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
