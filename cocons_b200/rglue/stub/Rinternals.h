/* STAND-IN DECLARATIONS ONLY.  R is not installed in the build image, so cocons_glue.c is
 * compiled against these declarations (the subset of R's C API it uses, with R's documented
 * prototypes) and linked, for the tests, with the miniature runtime in tests/rmock/rmock.c.
 * A real build uses R's own <Rinternals.h>. */
#ifndef COCONS_STUB_RINTERNALS_H
#define COCONS_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
#define STRSXP 16
#define EXTPTRSXP 22
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
extern double R_NaReal;
double* REAL(SEXP);
int* INTEGER(SEXP);
int TYPEOF(SEXP);
R_xlen_t XLENGTH(SEXP);
int LENGTH(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char* CHAR(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_coerceVector(SEXP, unsigned int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
int Rf_isMatrix(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_isNull(SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
void Rf_warning(const char*, ...);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_mkString(const char*);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
