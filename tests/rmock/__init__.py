"""TEST SCAFFOLDING: build cocons_b200/rglue/cocons_glue.c against the miniature R runtime in rmock.c and drive it
the way R would - `.Call(name, ...)` by REGISTERED name and arity, R objects in, R objects out, R errors as Python
exceptions - with R's gctorture-style checks always on (see rmock.c).  Python values map to R values like this:

    dict (str -> value)     named list            float / int scalar     length-1 double / integer vector
    2-D float / int array   double / integer matrix (column-major, `dim` attribute)
    1-D float / int array   double / integer vector
    None                    NULL                  ExtPtr                 an external pointer returned earlier
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GLUE = os.path.join(ROOT, "cocons_b200", "rglue")
LIBDIR = os.path.join(ROOT, "cocons_b200")

NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, EXTPTRSXP = 0, 10, 13, 14, 16, 19, 22


class RError(RuntimeError):
    """an R condition raised by the glue through Rf_error()"""


class RCheckError(AssertionError):
    """the glue broke a rule of R's C API (missing PROTECT, stack imbalance, leaked buffer)"""


class ExtPtr:
    def __init__(self, sexp):
        self.sexp = sexp


class RMock:
    def __init__(self, workdir, glue_source=None, library=None):
        """glue_source: another C file registering its routines through R_init_cocons (the self-test of the checks);
        default: the product's cocons_b200/rglue/cocons_glue.c.
        library: path of the shared object providing the C ABI; default libcocons_b200.so (the product).  The CPU suite
        also links the glue against the HOST BUILD of the library's sources (tests/host_emul) to run the .Call sequences
        end to end without a GPU."""
        workdir = str(workdir)
        alloc_h = os.path.join(workdir, "rmock_alloc.h")
        with open(alloc_h, "w") as f:
            f.write("#include <stdlib.h>\nvoid* rmock_malloc(size_t);\nvoid rmock_free(void*);\n"
                    "#define malloc rmock_malloc\n#define free rmock_free\n")
        inc = ["-I" + os.path.join(GLUE, "stub")]
        if os.environ.get("COCONS_EMUL_SANITIZE"):  # tools/emul_memcheck.sh: the glue's marshalling under ASan + UBSan
            inc += ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
        env = dict(os.environ)
        env.pop("LD_PRELOAD", None)  # the compiler itself does not run under the preloaded sanitizer runtime
        glue_o, mock_o = os.path.join(workdir, "glue.o"), os.path.join(workdir, "rmock.o")
        so = os.path.join(workdir, "cocons_rmock.so")
        subprocess.check_call(["gcc", "-O1", "-g", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter", "-c"] + inc +
                              ["-include", alloc_h, glue_source or os.path.join(GLUE, "cocons_glue.c"), "-o", glue_o], env=env)
        subprocess.check_call(["gcc", "-O1", "-g", "-fPIC", "-c"] + inc + [os.path.join(HERE, "rmock.c"), "-o", mock_o],
                              env=env)
        if library is None:
            link = ["-L" + LIBDIR, "-lcocons_b200", "-Wl,-rpath," + LIBDIR]
        else:
            link = [library, "-Wl,-rpath," + os.path.dirname(library)]
        subprocess.check_call(["gcc", "-shared", glue_o, mock_o] + link + ["-o", so], env=env)
        L = self.lib = ctypes.CDLL(so)
        vp, ci, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_long
        for name, res, args in (
                ("rmock_init", None, []), ("rmock_n_routines", ci, []), ("rmock_routine_name", ctypes.c_char_p, [ci]),
                ("rmock_routine_nargs", ci, [ci]), ("rmock_dynamic_symbols", ci, []), ("rmock_set_torture", None, [ci]),
                ("rmock_new", vp, [ci, cl]), ("rmock_nil", vp, []), ("rmock_dataptr", vp, [vp]), ("rmock_type", ci, [vp]),
                ("rmock_len", cl, [vp]), ("rmock_set_dim", None, [vp, ci, ci]),
                ("rmock_get_dim", ci, [vp, ctypes.POINTER(ci), ctypes.POINTER(ci)]),
                ("rmock_set_name", None, [vp, cl, ctypes.c_char_p]), ("rmock_set_elt", None, [vp, cl, vp]),
                ("rmock_get_elt", vp, [vp, cl]), ("rmock_last_error", ctypes.c_char_p, []), ("rmock_faults", ci, []),
                ("rmock_fault_msg", ctypes.c_char_p, []), ("rmock_live_mallocs", cl, []),
                ("rmock_protect_depth", ci, []), ("rmock_reset_faults", None, []),
                ("rmock_call", vp, [ctypes.c_char_p, ci, ctypes.POINTER(vp), ctypes.POINTER(ci)]),
                ("rmock_release_all", None, []), ("rmock_extptr", vp, [vp])):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        L.rmock_init()

    # ---- registration table, as R sees it after R_init_cocons -------------------------------------------------
    def routines(self):
        L = self.lib
        return {L.rmock_routine_name(i).decode(): L.rmock_routine_nargs(i) for i in range(L.rmock_n_routines())}

    def dynamic_symbols(self):
        return self.lib.rmock_dynamic_symbols()

    # ---- Python -> R -------------------------------------------------------------------------------------------
    def to_r(self, v):
        L = self.lib
        if v is None:
            return L.rmock_nil()
        if isinstance(v, ExtPtr):
            return v.sexp
        if isinstance(v, dict):
            s = L.rmock_new(VECSXP, len(v))
            for i, (k, x) in enumerate(v.items()):
                L.rmock_set_name(s, i, str(k).encode())
                L.rmock_set_elt(s, i, self.to_r(x))
            return s
        if isinstance(v, (list, tuple)) and any(isinstance(x, (dict, np.ndarray)) for x in v):
            s = L.rmock_new(VECSXP, len(v))  # unnamed list
            for i, x in enumerate(v):
                L.rmock_set_elt(s, i, self.to_r(x))
            return s
        a = np.asarray(v)
        if a.dtype.kind in "iub":
            a, typ = np.asfortranarray(a, dtype=np.int32), INTSXP
        else:
            a, typ = np.asfortranarray(a, dtype=np.float64), REALSXP
        s = L.rmock_new(typ, a.size)
        if a.size:
            ctypes.memmove(L.rmock_dataptr(s), a.ctypes.data, a.nbytes)
        if a.ndim == 2:
            L.rmock_set_dim(s, a.shape[0], a.shape[1])
        return s

    # ---- R -> Python -------------------------------------------------------------------------------------------
    def from_r(self, s):
        L = self.lib
        typ, n = L.rmock_type(s), L.rmock_len(s)
        if typ == NILSXP:
            return None
        if typ == EXTPTRSXP:
            return ExtPtr(s)
        if typ == VECSXP:
            return [self.from_r(L.rmock_get_elt(s, i)) for i in range(n)]
        dt = np.float64 if typ == REALSXP else np.int32
        out = np.empty(n, dtype=dt)
        if n:
            ctypes.memmove(out.ctypes.data, L.rmock_dataptr(s), out.nbytes)
        nr, nc = ctypes.c_int(), ctypes.c_int()
        if L.rmock_get_dim(s, ctypes.byref(nr), ctypes.byref(nc)):
            return out.reshape((nr.value, nc.value), order="F")
        return out

    def call(self, name, *args):
        """.Call(name, ...) with every check of rmock.c; raises RError for an R error, RCheckError for an API misuse"""
        L = self.lib
        L.rmock_reset_faults()
        mallocs0 = L.rmock_live_mallocs()
        sexps = (ctypes.c_void_p * max(len(args), 1))(*[self.to_r(a) for a in args])
        status = ctypes.c_int()
        out = L.rmock_call(name.encode(), len(args), sexps, ctypes.byref(status))
        if L.rmock_faults():
            raise RCheckError("%s: %s" % (name, L.rmock_fault_msg().decode()))
        if L.rmock_live_mallocs() != mallocs0:
            raise RCheckError("%s: %d malloc'd buffer(s) not freed" % (name, L.rmock_live_mallocs() - mallocs0))
        if status.value == 1:
            raise RError(L.rmock_last_error().decode())
        if status.value != 0:
            raise RCheckError(L.rmock_last_error().decode())
        return self.from_r(out)

    def release_all(self):
        self.lib.rmock_release_all()
