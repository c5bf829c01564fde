"""TEST SCAFFOLDING: compile every source of libcocons_b200.so (cocons_b200/csrc/*.cu) for the HOST against the emulation shim in
this directory and load the result with ctypes (see cuda_runtime.h here for the execution model).

The only edits made to the shipped sources are mechanical: every `kernel<<<grid, block, smem, stream>>>(args);`
becomes `emul::launch(grid, block, has_barrier, smem, [&] { kernel(args); });` (g++ cannot parse the chevrons), where
has_barrier says whether the kernel's body contains a barrier; `extern __shared__ T x[];` becomes a pointer to the
launch's dynamic shared memory; inline PTX is handled by the two rules of rewrite_ptx(); a cooperative launch becomes
a launch of ONE block (rewrite_cooperative_launches()); `#include "x"` lines are made
absolute because the rewritten text is compiled from a scratch directory."""
import ctypes
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cocons_b200", "csrc")
SOURCES = ("assembly.cu", "taper.cu", "solve.cu", "chol.cu", "dist.cu", "capi.cu")


def _matching(text, start, open_ch, close_ch):
    """index just past the bracket that closes text[start] (which must be open_ch)"""
    assert text[start] == open_ch
    depth = 0
    for k in range(start, len(text)):
        if text[k] == open_ch:
            depth += 1
        elif text[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k + 1
    raise ValueError("unbalanced %s" % open_ch)


def _split_top_level(s, angle=False):
    """split at the commas outside brackets (angle: template argument lists count as brackets too)"""
    parts, depth, cur = [], 0, ""
    for k, ch in enumerate(s):
        arrow = ch == ">" and k > 0 and s[k - 1] == "-"  # `->` is not a bracket
        if ch in "([{" or (angle and ch == "<"):
            depth += 1
        elif ch in ")]}" or (angle and ch == ">" and not arrow):
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def kernels_with_barriers(text):
    """names of the __global__ functions whose body calls __syncthreads() / grid.sync()"""
    out = {}
    for m in re.finditer(r"__global__\s+void\s+", text):
        k = m.end()
        if text.startswith("__launch_bounds__", k):
            k = _matching(text, text.index("(", k), "(", ")")
        name = re.match(r"\s*(\w+)\s*\(", text[k:])
        params_end = _matching(text, k + name.end() - 1, "(", ")")
        body_start = text.index("{", params_end)
        body = text[body_start:_matching(text, body_start, "{", "}")]
        out[name.group(1)] = "__syncthreads" in body or "grid.sync" in body
    return out


def rewrite_launches(text):
    text = text.replace("\\\n", " ")  # launches inside multi-line macros
    barriers = kernels_with_barriers(text)
    out, pos, count = "", 0, 0
    for m in re.finditer(r"(\w+)\s*(<[^<>;(){}]*>)?\s*<<<", text):
        if m.start() < pos:
            continue
        cfg_end = text.index(">>>", m.end())
        cfg = _split_top_level(text[m.end():cfg_end], angle=True)
        assert 2 <= len(cfg) <= 4, "expected <<<grid, block[, smem[, stream]]>>>: %r" % (cfg,)
        cfg += ["0", "nullptr"][len(cfg) - 2:]
        args_start = cfg_end + 3
        while text[args_start].isspace():
            args_start += 1
        args_end = _matching(text, args_start, "(", ")")
        name, targs = m.group(1), m.group(2) or ""
        assert name in barriers, "launch of an unknown kernel %s" % name
        out += text[pos:m.start()]
        out += "emul::launch(%s, %s, %s, %s, [&] { %s%s%s; })" % (
            cfg[0], cfg[1], "true" if barriers[name] else "false", cfg[2], name, targs, text[args_start:args_end])
        pos = args_end
        count += 1
    return out + text[pos:], count, barriers


def rewrite_cooperative_launches(text):
    """`void* args[] = {(void*)&a, (void*)&b, ...}; ... cudaLaunchCooperativeKernel((void*)k<T>, grid, block, args, smem,
    st);` -> the kernel on a grid of ONE block (blocks run one after another here, and a cooperative kernel needs all
    of its blocks resident: with one block grid.sync() is the block barrier; the kernels stride over gridDim.x, so any
    grid size computes the same thing)."""
    out, pos = "", 0
    for m in re.finditer(r"cudaLaunchCooperativeKernel\s*\(", text):
        end = _matching(text, m.end() - 1, "(", ")")
        parts = _split_top_level(text[m.end():end - 1], angle=True)
        assert len(parts) == 6, parts
        kernel = re.sub(r"^\(void\s*\*\)\s*", "", parts[0])
        decl = list(re.finditer(r"void\s*\*\s*%s\s*\[\]\s*=\s*\{([^}]*)\}" % re.escape(parts[3]), text[:m.start()]))[-1]
        names = [re.sub(r"^\(void\s*\*\)\s*&\s*", "", a.strip()) for a in decl.group(1).split(",")]
        out += text[pos:m.start()] + "emul::launch(dim3(1), %s, true, %s, [&] { %s(%s); })" % (
            parts[2], parts[4], kernel, ", ".join(names))
        pos = end
    return out + text[pos:]


def rewrite_dynamic_shared(text):
    """`extern __shared__ [__align__(n)] T name[];` -> a pointer to the launch's dynamic shared memory"""
    return re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];",
                  r"\1* \2 = static_cast<\1*>(emul::ctx.dyn_smem);", text)


def rewrite_ptx(text):
    """Inline PTX cannot be assembled for the host.  Two mechanical rules:
    (1) a `__device__ __forceinline__` helper whose body is an asm statement (mma.sync, mbarrier.*, cp.async.bulk,
        ld.acquire ...) gets the body `return emul::ptx_<name>(<its parameters>);` - the stand-ins live in
        cuda_runtime.h of this directory and restate the instruction's documented semantics;
    (2) a free-standing `asm volatile("fence..." / "prefetch..." ...)` statement inside a kernel has no functional
        effect in a sequentially consistent, single-threaded execution and becomes `(void)0;` - anything else is an
        error, so that no instruction is dropped silently.  (`asm volatile("" ::: "memory")` is left alone.)"""
    out, pos, helpers = "", 0, []
    for m in re.finditer(r"__device__\s+__forceinline__\s+([\w:\s\*&]+?)\s*\b(\w+)\s*\(", text):
        if m.start() < pos:
            continue
        params_end = _matching(text, m.end() - 1, "(", ")")
        k = params_end
        while text[k].isspace():
            k += 1
        if text[k] != "{":
            continue
        body_end = _matching(text, k, "{", "}")
        body = text[k:body_end]
        if not re.search(r"\basm\b", body):
            continue
        assert len(body) < 900, "an asm helper with a long body: %s" % m.group(2)
        params = [q.strip() for q in _split_top_level(text[m.end():params_end - 1]) if q.strip()]
        names = [re.search(r"(\w+)\s*$", q).group(1) for q in params]
        out += text[pos:k] + "{ return emul::ptx_%s(%s); }" % (m.group(2), ", ".join(names))
        pos = body_end
        helpers.append(m.group(2))
    text = out + text[pos:]
    out, pos = "", 0
    for m in re.finditer(r"\basm\s+volatile\s*\(", text):
        end = _matching(text, m.end() - 1, "(", ")")
        while text[end].isspace():
            end += 1
        assert text[end] == ";"
        instr = re.match(r'\s*"([^"]*)"', text[m.end():]).group(1).strip()
        if instr == "":
            continue
        assert instr.startswith("fence.") or instr.startswith("prefetch."), "unexpected PTX in a kernel body: %s" % instr
        out += text[pos:m.start()] + "(void)0;"
        pos = end + 1
    return out + text[pos:], helpers


def _absolute_includes(text):
    def fix(m):
        return '#include "%s"' % os.path.normpath(os.path.join(CSRC, m.group(1)))
    return re.sub(r'#include\s+"([^"]+)"', fix, text)


def _compiler_env():
    """the compiler itself must not run under a preloaded sanitizer runtime (tools/emul_memcheck.sh preloads ASan)"""
    env = dict(os.environ)
    env.pop("LD_PRELOAD", None)
    return env


def _rewrite_sources(workdir):
    info, launches, helpers = {}, 0, []
    for src in SOURCES:
        text, names = rewrite_ptx(rewrite_dynamic_shared(rewrite_cooperative_launches(
            open(os.path.join(CSRC, src)).read())))
        text, count, barriers = rewrite_launches(text)
        info.update(barriers)
        helpers += names
        launches += count
        with open(os.path.join(workdir, src.replace(".cu", "_emul.inc")), "w") as f:
            f.write(_absolute_includes(text))
    return info, launches, helpers


INC_DEFINES = ['-DASSEMBLY_INC="assembly_emul.inc"', '-DTAPER_INC="taper_emul.inc"', '-DSOLVE_INC="solve_emul.inc"',
               '-DCHOL_INC="chol_emul.inc"', '-DDIST_INC="dist_emul.inc"', '-DCAPI_INC="capi_emul.inc"']


def build_racecheck(workdir):
    """-> path of the race-check executable: the host build with every CUDA thread a ThreadSanitizer fiber
    (racecheck_main.cpp in this directory)"""
    workdir = str(workdir)
    _rewrite_sources(workdir)
    exe = os.path.join(workdir, "racecheck")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fsanitize=thread", "-Wno-tsan",
                           "-DCOCONS_EMUL_TSAN", "-I" + HERE, "-I" + workdir] + INC_DEFINES +
                          [os.path.join(HERE, "racecheck_main.cpp"), "-o", exe], env=_compiler_env())
    return exe


def build(workdir):
    """-> (ctypes library, {kernel name: has_barrier}, number of rewritten launches)"""
    workdir = str(workdir)
    info, launches, helpers = _rewrite_sources(workdir)
    so = os.path.join(workdir, "libcocons_host_emul.so")
    # COCONS_EMUL_SANITIZE=1: AddressSanitizer + UBSan over the kernels (heap = "device" memory, globals = "shared"
    # memory; stack instrumentation off because the fibers switch stacks) - the memcheck run of tools/emul_memcheck.sh
    san = (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "--param", "asan-stack=0", "-g",
            "-fno-omit-frame-pointer"] if os.environ.get("COCONS_EMUL_SANITIZE") else [])
    subprocess.check_call([
        "g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC"] + san + [
        "-I" + HERE, "-I" + workdir] + INC_DEFINES + ["-Wl,-Bsymbolic",
        os.path.join(HERE, "driver.cpp"), "-o", so], env=_compiler_env())
    lib = ctypes.CDLL(so)
    d, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.emu_launches.restype = ctypes.c_long
    lib.emu_barrier_launches.restype = ctypes.c_long
    lib.emu_cov_square.argtypes = [i32, i64, i64, d, d, d, d, d]
    lib.emu_cov_pred.argtypes = [i64, i64, i64, d, d, d, d, d, d, d]
    lib.emu_morton.argtypes = [i64, d, d]
    lib.emu_morton.restype = None
    lib.emu_ctx_lower.argtypes = [i32, i64, i64, i64, d, d, d, d, d, d]
    lib.emu_dist_slabs.argtypes = [i32, i64, i64, i64, d, d, d, d, d, i32, i32, i64, i32, d]
    lib.emu_taper_entries.argtypes = [i64, i64, i64, d, d, d, d, d, d, d, d, i64, d]
    lib.emu_taper_lower.argtypes = [i64, i64, i64, d, d, d, d, d, d, i64, d, d, d]
    lib.emu_forward_solve.argtypes = [i64, d, d, d, i32, i32]
    lib.emu_logdet.argtypes = [d, i64, i64]
    lib.emu_logdet.restype = ctypes.c_double
    lib.emu_gram.argtypes = [d, i64, i64, i32, d]
    lib.emu_gram.restype = None
    lib.emu_chol_factor.argtypes = [i64, d, d]
    lib.emu_potrf_tile.argtypes = [d, i64, d, i32]
    lib.emu_potrf_tile_sweep.argtypes = [d, i64, d, i32]
    lib.emu_gemm_nt.argtypes = [i32, i64, i64, i64, d, i64, d, i64, d, i64, i32]
    lib.emu_gemm_nt.restype = None
    lib.emu_tile_count.argtypes = [i32, i32, i32, i32]
    lib.emu_tile_count.restype = i64
    lib.emu_tile_decode_check.argtypes = [i32, ctypes.POINTER(ctypes.c_long)]
    lib.emu_tile_decode_check.restype = ctypes.c_long
    lib.emu_last_error.restype = ctypes.c_char_p
    lib.emu_set_concurrent_blocks.argtypes = [i32]
    lib.ptx_helpers = sorted(helpers)
    return lib, info, launches
