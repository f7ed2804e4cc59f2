/* TEST-ONLY stand-in for R's headers: the declarations r/src/rshim.c needs, implemented by
 * tests/rstub/fake_r.c so that the shim can be compiled AND executed on a machine without R
 * (tests/test_rshim_syntax.py, tests/test_rshim_exec.py).  Not used by any product build. */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stddef.h>
#include <stdio.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef void *(*DL_FUNC)(void);
typedef struct _DllInfo DllInfo;
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
extern SEXP R_NamesSymbol, R_NilValue, R_DimNamesSymbol, R_DimSymbol;
extern double R_NaN;
enum { INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19 };
SEXP Rf_getAttrib(SEXP, SEXP); SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
R_xlen_t XLENGTH(SEXP); const char *CHAR(SEXP); SEXP STRING_ELT(SEXP, R_xlen_t); SEXP VECTOR_ELT(SEXP, R_xlen_t);
double Rf_asReal(SEXP); int Rf_asInteger(SEXP); int Rf_asLogical(SEXP); int Rf_isNull(SEXP); int Rf_isMatrix(SEXP);
int Rf_nrows(SEXP); int Rf_ncols(SEXP);
double *REAL(SEXP); int *INTEGER(SEXP);
SEXP Rf_allocVector(unsigned, R_xlen_t); SEXP Rf_allocMatrix(unsigned, int, int); SEXP Rf_ScalarInteger(int);
SEXP Rf_mkChar(const char *); void SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char *, ...) __attribute__((noreturn));
char *R_alloc(size_t, int);
int R_registerRoutines(DllInfo *, const void *, const R_CallMethodDef *, const void *, const void *);
int R_useDynamicSymbols(DllInfo *, int);
#define FALSE 0
#endif
