/* minimal stub of the R C API, ONLY to compile-check rpkg/src/shim.c where R is not installed */
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef void* (*DL_FUNC)(void);
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define REALSXP 14
#define VECSXP 19
int TYPEOF(SEXP);
void R_CheckUserInterrupt(void);
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
int* INTEGER(SEXP); double* REAL(SEXP); R_xlen_t XLENGTH(SEXP); SEXP VECTOR_ELT(SEXP, R_xlen_t); SEXP STRING_ELT(SEXP, R_xlen_t);
const char* CHAR(SEXP); SEXP Rf_getAttrib(SEXP, SEXP); int Rf_asInteger(SEXP); double Rf_asReal(SEXP);
SEXP Rf_allocMatrix(int, int, int); SEXP Rf_allocVector(int, R_xlen_t); SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...) __attribute__((noreturn));
void* R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP); SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, void (*)(SEXP), Rboolean);
char* R_alloc(size_t, int);
