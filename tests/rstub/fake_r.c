/* TEST-ONLY miniature R runtime: enough of R's C API to EXECUTE r/src/rshim.c on a machine without R
 * (tests/test_rshim_exec.py).  SEXPs live on a heap of tagged vectors with names / dim / dimnames
 * attributes; PROTECT keeps a counter so that a test can assert the shim's PROTECT / UNPROTECT balance;
 * Rf_error long-jumps to the frame set by fake_r_try(), like R's own error handling does, so the
 * shim's "release native resources, then Rf_error" discipline is exercised for real.  Not a product
 * file: nothing here is linked into libsoundgen_b200.so. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "R.h"

struct SEXPREC {
  int type;              /* 0 = NULL */
  R_xlen_t len;
  void *data;            /* double / int / SEXP* / char* */
  SEXP names, dim, dimnames;
};
static struct SEXPREC nil_rec, names_sym, dim_sym, dimnames_sym;
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_sym, R_DimNamesSymbol = &dimnames_sym, R_DimSymbol = &dim_sym;
double R_NaN;
static int protect_depth = 0, protect_max = 0;
static jmp_buf *err_jmp = NULL;
static char err_msg[1024];
static void *ralloc_list[4096];
static int ralloc_n = 0;

static SEXP new_sexp(int type, R_xlen_t len, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = type; s->len = len;
  s->data = calloc((size_t)(len > 0 ? len : 1), elt);
  s->names = s->dim = s->dimnames = R_NilValue;
  return s;
}
SEXP Rf_allocVector(unsigned type, R_xlen_t n) {
  switch (type) {
    case REALSXP: return new_sexp(REALSXP, n, sizeof(double));
    case INTSXP: return new_sexp(INTSXP, n, sizeof(int));
    case STRSXP: case VECSXP: {
      SEXP s = new_sexp((int)type, n, sizeof(SEXP));
      for (R_xlen_t i = 0; i < n; i++) ((SEXP *)s->data)[i] = R_NilValue;
      return s;
    }
  }
  Rf_error("fake_r: unsupported type %u", type);
}
SEXP Rf_allocMatrix(unsigned type, int nr, int nc) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc);
  s->dim = new_sexp(INTSXP, 2, sizeof(int));
  ((int *)s->dim->data)[0] = nr; ((int *)s->dim->data)[1] = nc;
  return s;
}
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); ((int *)s->data)[0] = v; return s; }
SEXP Rf_mkChar(const char *c) {
  SEXP s = new_sexp(9 /* CHARSXP */, (R_xlen_t)strlen(c), 1);
  free(s->data);
  s->data = strdup(c);
  return s;
}
R_xlen_t XLENGTH(SEXP s) { return s->type ? s->len : 0; }
const char *CHAR(SEXP s) { return (const char *)s->data; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { return ((SEXP *)s->data)[i]; }
void SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { ((SEXP *)s->data)[i] = v; }
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) { ((SEXP *)s->data)[i] = v; }
double *REAL(SEXP s) { if (s->type != REALSXP) Rf_error("fake_r: REAL() of a non-double vector"); return (double *)s->data; }
int *INTEGER(SEXP s) { if (s->type != INTSXP) Rf_error("fake_r: INTEGER() of a non-integer vector"); return (int *)s->data; }
int Rf_isNull(SEXP s) { return s == R_NilValue || s->type == 0; }
int Rf_isMatrix(SEXP s) { return !Rf_isNull(s) && !Rf_isNull(s->dim) && s->dim->len == 2; }
int Rf_nrows(SEXP s) { return Rf_isMatrix(s) ? ((int *)s->dim->data)[0] : (int)s->len; }
int Rf_ncols(SEXP s) { return Rf_isMatrix(s) ? ((int *)s->dim->data)[1] : 1; }
double Rf_asReal(SEXP s) {
  if (Rf_isNull(s) || s->len < 1) return R_NaN;
  return s->type == REALSXP ? ((double *)s->data)[0] : (double)((int *)s->data)[0];
}
int Rf_asInteger(SEXP s) { return (int)Rf_asReal(s); }
int Rf_asLogical(SEXP s) { return Rf_asReal(s) != 0; }
SEXP Rf_getAttrib(SEXP s, SEXP which) {
  if (which == R_NamesSymbol) return s->names;
  if (which == R_DimNamesSymbol) return s->dimnames;
  if (which == R_DimSymbol) return s->dim;
  return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP which, SEXP v) {
  if (which == R_NamesSymbol) s->names = v;
  else if (which == R_DimNamesSymbol) s->dimnames = v;
  else if (which == R_DimSymbol) s->dim = v;
  return v;
}
SEXP Rf_protect(SEXP s) { protect_depth++; if (protect_depth > protect_max) protect_max = protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; }
char *R_alloc(size_t n, int size) {
  void *p = calloc(n ? n : 1, (size_t)size);
  if (ralloc_n < 4096) ralloc_list[ralloc_n++] = p;
  return (char *)p;
}
void Rf_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_msg, sizeof err_msg, fmt, ap);
  va_end(ap);
  if (!err_jmp) { fprintf(stderr, "fake_r: uncaught error: %s\n", err_msg); abort(); }
  longjmp(*err_jmp, 1);
}
int R_registerRoutines(DllInfo *d, const void *c, const R_CallMethodDef *call, const void *f, const void *e) {
  (void)d; (void)c; (void)f; (void)e;
  int n = 0;
  while (call[n].name) n++;
  return n;
}
int R_useDynamicSymbols(DllInfo *d, int v) { (void)d; (void)v; return 0; }

/* ---- test-side helpers ---- */
void fake_r_init(void) { R_NaN = 0.0 / 0.0; nil_rec.type = 0; }
int fake_r_protect_depth(void) { return protect_depth; }
int fake_r_protect_max(void) { return protect_max; }
const char *fake_r_last_error(void) { return err_msg; }
void fake_r_end_call(void) {            /* what R does when .Call returns: R_alloc memory is released */
  for (int i = 0; i < ralloc_n; i++) free(ralloc_list[i]);
  ralloc_n = 0;
}
/* Runs f(arg) like a top-level R call: returns 0, or 1 if Rf_error was raised (the protect stack is then
 * unwound to its depth at entry, as R does). */
int fake_r_try(void (*f)(void *), void *arg) {
  jmp_buf jb, *prev = err_jmp;
  int depth = protect_depth;
  err_jmp = &jb;
  int failed = setjmp(jb);
  if (!failed) f(arg);
  else protect_depth = depth;
  err_jmp = prev;
  fake_r_end_call();
  return failed;
}
SEXP fake_r_real(const double *v, R_xlen_t n) { SEXP s = Rf_allocVector(REALSXP, n); if (n) memcpy(s->data, v, sizeof(double) * (size_t)n); return s; }
SEXP fake_r_int(const int *v, R_xlen_t n) { SEXP s = Rf_allocVector(INTSXP, n); if (n) memcpy(s->data, v, sizeof(int) * (size_t)n); return s; }
SEXP fake_r_matrix(const double *colmajor, int nr, int nc) { SEXP s = Rf_allocMatrix(REALSXP, nr, nc); memcpy(s->data, colmajor, sizeof(double) * (size_t)nr * nc); return s; }
SEXP fake_r_named_list(const char **names, const double *vals, int n) {
  SEXP l = Rf_allocVector(VECSXP, n), nm = Rf_allocVector(STRSXP, n);
  for (int i = 0; i < n; i++) { SET_VECTOR_ELT(l, i, fake_r_real(&vals[i], 1)); SET_STRING_ELT(nm, i, Rf_mkChar(names[i])); }
  Rf_setAttrib(l, R_NamesSymbol, nm);
  return l;
}
SEXP fake_r_list_get(SEXP l, const char *name) {
  for (R_xlen_t i = 0; i < l->len; i++) if (!strcmp(CHAR(STRING_ELT(l->names, i)), name)) return VECTOR_ELT(l, i);
  return R_NilValue;
}
