/* TEST SCAFFOLDING - a miniature of R's C API, just enough to LOAD AND RUN cocons_b200/rglue/cocons_glue.c
 * without R (R is not installed in this image): the subset declared in cocons_b200/rglue/stub/, with R's documented
 * semantics, plus the checks R itself only makes under gctorture / --use-valgrind:
 *   - objects are garbage-collected AT EVERY ALLOCATION: anything not reachable from the PROTECT stack, the
 *     arguments of the running .Call or the harness's own handles is poisoned (data overwritten, flagged), and any
 *     later access through the API is recorded as a fault - a missing PROTECT shows up deterministically;
 *   - the PROTECT stack must be balanced when a .Call returns (R: "stack imbalance in .Call");
 *   - Rf_error() unwinds to the harness like R's longjmp to top level, after restoring the PROTECT depth;
 *   - the glue is compiled with malloc / free redirected to counting wrappers, so a buffer that an error path
 *     forgets shows up as a leak.
 * Python drives it through ctypes (tests/rmock/__init__.py); nothing under cocons_b200/ links this. */
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LGLSXP 10
#define NILSXP 0
#define CHARSXP 9
#define EXTPTRSXP 22

struct SEXPREC {
  int type;
  R_xlen_t len;
  void* data;      /* double[], int[], SEXP[] (VECSXP / STRSXP), char[] (CHARSXP) */
  SEXP names, dim; /* the two attributes the glue reads or R sets on its results */
  void* ext;
  R_CFinalizer_t fin;
  int poisoned, mark, preserved;
  struct SEXPREC* next;
};

static struct SEXPREC nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL, 0, 0, 1, NULL};
static struct SEXPREC names_sym = {1, 0, NULL, NULL, NULL, NULL, NULL, 0, 0, 1, NULL};
static struct SEXPREC dim_sym = {1, 0, NULL, NULL, NULL, NULL, NULL, 0, 0, 1, NULL};
SEXP R_NilValue = &nil_obj, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;
double R_NaReal;

static SEXP all_objects = NULL;
static SEXP protect_stack[10000];
static int protect_top = 0;
static SEXP call_args[16];
static int call_nargs = 0;
static int faults = 0;
static char fault_msg[512], error_msg[1024];
static jmp_buf top_level;
static int in_call = 0;
static long live_mallocs = 0;
static int torture = 1;

static void fault(const char* what) {
  if (!faults) snprintf(fault_msg, sizeof fault_msg, "%s", what);
  ++faults;
}

static SEXP checked(SEXP s, const char* who) {
  if (!s) {
    fault("NULL SEXP passed to the R API");
    return R_NilValue;
  }
  if (s->poisoned) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s on an object that was not protected across an allocation (type %d, length %ld)", who,
             s->type, (long)s->len);
    fault(buf);
  }
  return s;
}

/* ---- collector: mark from the roots, poison the rest ---------------------------------------- */
static void mark(SEXP s) {
  if (!s || s == R_NilValue || s->mark) return;
  s->mark = 1;
  mark(s->names);
  mark(s->dim);
  if ((s->type == VECSXP || s->type == STRSXP) && s->data)
    for (R_xlen_t i = 0; i < s->len; ++i) mark(((SEXP*)s->data)[i]);
}

static void collect(void) {
  for (SEXP o = all_objects; o; o = o->next) o->mark = 0;
  for (SEXP o = all_objects; o; o = o->next)
    if (o->preserved) mark(o);
  for (int i = 0; i < protect_top; ++i) mark(protect_stack[i]);
  for (int i = 0; i < call_nargs; ++i) mark(call_args[i]);
  for (SEXP o = all_objects; o; o = o->next)
    if (!o->mark && !o->poisoned) {
      if (o->type == EXTPTRSXP && o->fin) { /* R runs the finalizer of an unreachable external pointer */
        R_CFinalizer_t f = o->fin;
        o->fin = NULL;
        f(o);
      }
      o->poisoned = 1;
      if (o->type == REALSXP || o->type == INTSXP || o->type == LGLSXP)
        memset(o->data, 0xFF, (size_t)o->len * (o->type == REALSXP ? sizeof(double) : sizeof(int)));
    }
}

static SEXP new_obj(int type, R_xlen_t len) {
  if (torture && in_call) collect();
  SEXP s = (SEXP)calloc(1, sizeof *s);
  s->type = type, s->len = len, s->names = R_NilValue, s->dim = R_NilValue;
  size_t bytes = 0;
  if (type == REALSXP) bytes = sizeof(double) * (size_t)len;
  if (type == INTSXP || type == LGLSXP) bytes = sizeof(int) * (size_t)len;
  if (type == VECSXP || type == STRSXP) bytes = sizeof(SEXP) * (size_t)len;
  if (type == CHARSXP) bytes = (size_t)len + 1;
  s->data = calloc(bytes ? bytes : 1, 1);
  if (type == VECSXP || type == STRSXP)
    for (R_xlen_t i = 0; i < len; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  if (type == REALSXP) /* R does not zero numeric vectors: make reliance on that visible */
    for (R_xlen_t i = 0; i < len; ++i) ((double*)s->data)[i] = -12345.678;
  s->next = all_objects;
  all_objects = s;
  return s;
}

/* ---- the API subset (prototypes in cocons_b200/rglue/stub/Rinternals.h) ------------------------ */
double* REAL(SEXP s) {
  s = checked(s, "REAL()");
  if (s->type != REALSXP) fault("REAL() on a non-double object");
  return (double*)s->data;
}
int* INTEGER(SEXP s) {
  s = checked(s, "INTEGER()");
  if (s->type != INTSXP && s->type != LGLSXP) fault("INTEGER() on a non-integer object");
  return (int*)s->data;
}
int TYPEOF(SEXP s) { return checked(s, "TYPEOF()")->type; }
R_xlen_t XLENGTH(SEXP s) { return checked(s, "XLENGTH()")->len; }
int LENGTH(SEXP s) { return (int)checked(s, "LENGTH()")->len; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) {
  s = checked(s, "VECTOR_ELT()");
  if (s->type != VECSXP || i < 0 || i >= s->len) {
    fault("VECTOR_ELT() out of range or not a list");
    return R_NilValue;
  }
  return ((SEXP*)s->data)[i];
}
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) {
  s = checked(s, "SET_VECTOR_ELT()");
  checked(v, "SET_VECTOR_ELT(value)");
  if (s->type != VECSXP || i < 0 || i >= s->len) {
    fault("SET_VECTOR_ELT() out of range or not a list");
    return v;
  }
  ((SEXP*)s->data)[i] = v;
  return v;
}
SEXP STRING_ELT(SEXP s, R_xlen_t i) {
  s = checked(s, "STRING_ELT()");
  if (s->type != STRSXP || i < 0 || i >= s->len) {
    fault("STRING_ELT() out of range or not a character vector");
    return R_NilValue;
  }
  return ((SEXP*)s->data)[i];
}
const char* CHAR(SEXP s) {
  s = checked(s, "CHAR()");
  return s->type == CHARSXP ? (const char*)s->data : "";
}
SEXP Rf_getAttrib(SEXP s, SEXP which) {
  s = checked(s, "getAttrib()");
  if (which == R_NamesSymbol) return s->names;
  if (which == R_DimSymbol) return s->dim;
  return R_NilValue;
}
SEXP Rf_allocVector(unsigned int type, R_xlen_t n) { return new_obj((int)type, n); }
SEXP Rf_protect(SEXP s) {
  checked(s, "PROTECT()");
  if (protect_top < 10000) protect_stack[protect_top++] = s;
  return s;
}
void Rf_unprotect(int n) {
  if (n > protect_top) {
    fault("UNPROTECT(): stack underflow");
    n = protect_top;
  }
  protect_top -= n;
}
SEXP Rf_allocMatrix(unsigned int type, int nr, int nc) {
  SEXP m = Rf_protect(new_obj((int)type, (R_xlen_t)nr * nc));
  SEXP d = new_obj(INTSXP, 2);
  ((int*)d->data)[0] = nr, ((int*)d->data)[1] = nc;
  m->dim = d;
  Rf_unprotect(1);
  return m;
}
SEXP Rf_coerceVector(SEXP s, unsigned int type) {
  s = checked(s, "coerceVector()");
  if ((unsigned)s->type == type) return s;
  if (!((s->type == INTSXP || s->type == LGLSXP || s->type == REALSXP) && (type == REALSXP || type == INTSXP))) {
    Rf_error("cannot coerce type %d to vector of type %u", s->type, type);
  }
  Rf_protect(s);
  SEXP out = new_obj((int)type, s->len); /* attributes (dim, names) are kept, as in R */
  Rf_unprotect(1);
  out->dim = s->dim, out->names = s->names;
  for (R_xlen_t i = 0; i < s->len; ++i) {
    if (type == REALSXP)
      ((double*)out->data)[i] = (double)((int*)s->data)[i];
    else
      ((int*)out->data)[i] = (int)((double*)s->data)[i];
  }
  return out;
}
int Rf_isMatrix(SEXP s) {
  s = checked(s, "isMatrix()");
  return s->dim != R_NilValue && s->dim->len == 2;
}
int Rf_nrows(SEXP s) { /* R: a vector without dim has length(s) rows and one column */
  s = checked(s, "nrows()");
  if (s->dim != R_NilValue) return ((int*)s->dim->data)[0];
  if (s->type == REALSXP || s->type == INTSXP || s->type == LGLSXP || s->type == VECSXP) return (int)s->len;
  Rf_error("object is not a matrix");
}
int Rf_ncols(SEXP s) {
  s = checked(s, "ncols()");
  if (s->dim != R_NilValue) return s->dim->len >= 2 ? ((int*)s->dim->data)[1] : 1;
  if (s->type == REALSXP || s->type == INTSXP || s->type == LGLSXP || s->type == VECSXP) return 1;
  Rf_error("object is not a matrix");
}
int Rf_asInteger(SEXP s) {
  s = checked(s, "asInteger()");
  if (s->len < 1) return -2147483647 - 1; /* NA_INTEGER */
  if (s->type == INTSXP || s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == REALSXP) return (int)((double*)s->data)[0];
  return -2147483647 - 1;
}
double Rf_asReal(SEXP s) {
  s = checked(s, "asReal()");
  if (s->len < 1) return R_NaReal;
  if (s->type == REALSXP) return ((double*)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) return (double)((int*)s->data)[0];
  return R_NaReal;
}
int Rf_isNull(SEXP s) { return checked(s, "isNull()")->type == NILSXP; }
void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_msg, sizeof error_msg, fmt, ap);
  va_end(ap);
  if (!in_call) {
    fprintf(stderr, "rmock: Rf_error outside a .Call: %s\n", error_msg);
    abort();
  }
  longjmp(top_level, 1);
}
void Rf_warning(const char* fmt, ...) { (void)fmt; }
SEXP Rf_ScalarReal(double v) {
  SEXP s = new_obj(REALSXP, 1);
  ((double*)s->data)[0] = v;
  return s;
}
SEXP Rf_ScalarInteger(int v) {
  SEXP s = new_obj(INTSXP, 1);
  ((int*)s->data)[0] = v;
  return s;
}
SEXP Rf_mkString(const char* str) {
  SEXP s = Rf_protect(new_obj(STRSXP, 1));
  SEXP c = new_obj(CHARSXP, (R_xlen_t)strlen(str));
  memcpy(c->data, str, strlen(str));
  ((SEXP*)s->data)[0] = c;
  Rf_unprotect(1);
  return s;
}
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  (void)tag, (void)prot;
  SEXP s = new_obj(EXTPTRSXP, 0);
  s->ext = p;
  return s;
}
void* R_ExternalPtrAddr(SEXP s) {
  s = checked(s, "R_ExternalPtrAddr()");
  if (s->type != EXTPTRSXP) {
    fault("R_ExternalPtrAddr() on a non-pointer object");
    return NULL;
  }
  return s->ext;
}
void R_ClearExternalPtr(SEXP s) { checked(s, "R_ClearExternalPtr()")->ext = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fin, Rboolean onexit) {
  (void)onexit;
  checked(s, "R_RegisterCFinalizerEx()")->fin = fin;
}

/* ---- registration ----------------------------------------------------------------------------- */
static const R_CallMethodDef* routines = NULL;
static int dynamic_symbols = -1;
int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)dll, (void)c, (void)f, (void)e;
  routines = call;
  return 1;
}
int R_useDynamicSymbols(DllInfo* dll, int value) {
  (void)dll;
  dynamic_symbols = value;
  return 1;
}

/* ---- counting allocator the glue is compiled against (-Dmalloc=rmock_malloc -Dfree=rmock_free) ---- */
void* rmock_malloc(size_t n) {
  void* p = (malloc)(n);
  if (p) ++live_mallocs;
  return p;
}
void rmock_free(void* p) {
  if (p) --live_mallocs;
  (free)(p);
}

/* ---- harness side (ctypes) ---------------------------------------------------------------------- */
void R_init_cocons(DllInfo* dll);

void rmock_init(void) {
  const union {
    unsigned long long u;
    double d;
  } na = {0x7FF00000000007A2ull}; /* R's NA_real_: a NaN with payload 1954 */
  R_NaReal = na.d;
  R_init_cocons(NULL);
}
int rmock_n_routines(void) {
  int n = 0;
  while (routines && routines[n].name) ++n;
  return n;
}
const char* rmock_routine_name(int i) { return routines[i].name; }
int rmock_routine_nargs(int i) { return routines[i].numArgs; }
int rmock_dynamic_symbols(void) { return dynamic_symbols; }
void rmock_set_torture(int on) { torture = on; }

/* objects made by the harness stay alive until rmock_release_all() */
SEXP rmock_new(int type, long len) {
  SEXP s = new_obj(type, len);
  s->preserved = 1;
  if (type == REALSXP) memset(s->data, 0, sizeof(double) * (size_t)len);
  return s;
}
SEXP rmock_nil(void) { return R_NilValue; }
SEXP rmock_extptr(void* p) {
  SEXP s = new_obj(EXTPTRSXP, 0);
  s->preserved = 1, s->ext = p;
  return s;
}
void* rmock_dataptr(SEXP s) { return s->data; }
int rmock_type(SEXP s) { return s->type; }
long rmock_len(SEXP s) { return (long)s->len; }
int rmock_poisoned(SEXP s) { return s->poisoned; }
void rmock_set_dim(SEXP s, int nr, int nc) {
  SEXP d = rmock_new(INTSXP, 2);
  ((int*)d->data)[0] = nr, ((int*)d->data)[1] = nc;
  s->dim = d;
}
int rmock_get_dim(SEXP s, int* nr, int* nc) {
  if (s->dim == R_NilValue || s->dim->len != 2) return 0;
  *nr = ((int*)s->dim->data)[0], *nc = ((int*)s->dim->data)[1];
  return 1;
}
void rmock_set_name(SEXP s, long i, const char* name) {
  if (s->names == R_NilValue) s->names = rmock_new(STRSXP, s->len);
  SEXP c = rmock_new(CHARSXP, (long)strlen(name));
  memcpy(c->data, name, strlen(name));
  ((SEXP*)s->names->data)[i] = c;
}
void rmock_set_elt(SEXP s, long i, SEXP v) { ((SEXP*)s->data)[i] = v; }
SEXP rmock_get_elt(SEXP s, long i) { return ((SEXP*)s->data)[i]; }
void rmock_preserve(SEXP s) { s->preserved = 1; }

const char* rmock_last_error(void) { return error_msg; }
int rmock_faults(void) { return faults; }
const char* rmock_fault_msg(void) { return fault_msg; }
long rmock_live_mallocs(void) { return live_mallocs; }
int rmock_protect_depth(void) { return protect_top; }
void rmock_reset_faults(void) { faults = 0, fault_msg[0] = 0; }

typedef SEXP (*F0)(void);
typedef SEXP (*F1)(SEXP);
typedef SEXP (*F2)(SEXP, SEXP);
typedef SEXP (*F3)(SEXP, SEXP, SEXP);
typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

/* .Call(name, ...): NULL + rmock_last_error() when the routine raised an R error, when the name is not registered
 * or when the number of arguments is not the registered one (R checks that too).  status: 0 ok, 1 R error,
 * 2 registration problem, 3 PROTECT stack imbalance on return */
SEXP rmock_call(const char* name, int nargs, SEXP* args, int* status) {
  int k = 0;
  *status = 2;
  while (routines && routines[k].name && strcmp(routines[k].name, name) != 0) ++k;
  if (!routines || !routines[k].name) {
    snprintf(error_msg, sizeof error_msg, "\"%s\" not available for .Call()", name);
    return NULL;
  }
  if (routines[k].numArgs != nargs || nargs > 8) {
    snprintf(error_msg, sizeof error_msg, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs,
             routines[k].numArgs, name);
    return NULL;
  }
  DL_FUNC fn = routines[k].fun;
  const int depth = protect_top;
  call_nargs = nargs;
  for (int i = 0; i < nargs; ++i) call_args[i] = args[i];
  SEXP a[8];
  for (int i = 0; i < 8; ++i) a[i] = i < nargs ? args[i] : R_NilValue;
  volatile SEXP out = NULL;
  in_call = 1;
  if (setjmp(top_level) == 0) {
    switch (nargs) {
      case 0: out = ((F0)fn)(); break;
      case 1: out = ((F1)fn)(a[0]); break;
      case 2: out = ((F2)fn)(a[0], a[1]); break;
      case 3: out = ((F3)fn)(a[0], a[1], a[2]); break;
      case 4: out = ((F4)fn)(a[0], a[1], a[2], a[3]); break;
      case 5: out = ((F5)fn)(a[0], a[1], a[2], a[3], a[4]); break;
      case 6: out = ((F6)fn)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
      case 7: out = ((F7)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
      default: out = ((F8)fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
    }
    *status = (protect_top == depth) ? 0 : 3;
    if (*status == 3)
      snprintf(error_msg, sizeof error_msg, "stack imbalance in '.Call', %d then %d", depth, protect_top);
  } else {
    *status = 1; /* R error: the PROTECT stack is unwound by the top-level handler */
    out = NULL;
  }
  protect_top = depth;
  in_call = 0, call_nargs = 0;
  if (out) {
    checked((SEXP)out, "the value returned by .Call");
    ((SEXP)out)->preserved = 1; /* the harness holds it now */
  }
  return (SEXP)out;
}

/* frees every object (finalizers of external pointers run first, as at R's exit) */
void rmock_release_all(void) {
  for (SEXP o = all_objects; o; o = o->next)
    if (o->type == EXTPTRSXP && o->fin && !o->poisoned) {
      R_CFinalizer_t f = o->fin;
      o->fin = NULL;
      f(o);
    }
  while (all_objects) {
    SEXP n = all_objects->next;
    (free)(all_objects->data);
    (free)(all_objects);
    all_objects = n;
  }
  protect_top = 0;
}
