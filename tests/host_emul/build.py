"""TEST SCAFFOLDING: compile cocons_b200/csrc/assembly.cu and taper.cu for the HOST against the emulation shim in
this directory and load the result with ctypes (see cuda_runtime.h here for the execution model).

The only edit made to the shipped sources is mechanical: every `kernel<<<grid, block, smem, stream>>>(args);`
becomes `emul::launch(grid, block, has_barrier, [&] { kernel(args); });` (g++ cannot parse the chevrons), where
has_barrier says whether the kernel's body contains __syncthreads(); `#include "x"` lines are made absolute
because the rewritten text is compiled from a scratch directory."""
import ctypes
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cocons_b200", "csrc")


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


def _split_top_level(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def kernels_with_barriers(text):
    """names of the __global__ functions whose body calls __syncthreads()"""
    out = {}
    for m in re.finditer(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(", text):
        params_end = _matching(text, m.end() - 1, "(", ")")
        body_start = text.index("{", params_end)
        body = text[body_start:_matching(text, body_start, "{", "}")]
        out[m.group(1)] = "__syncthreads" in body
    return out


def rewrite_launches(text):
    text = text.replace("\\\n", " ")  # launches inside multi-line macros
    barriers = kernels_with_barriers(text)
    out, pos, count = "", 0, 0
    for m in re.finditer(r"(\w+)\s*(<\s*\w+\s*>)?\s*<<<", text):
        if m.start() < pos:
            continue
        cfg_end = text.index(">>>", m.end())
        cfg = _split_top_level(text[m.end():cfg_end])
        assert len(cfg) == 4, "expected <<<grid, block, smem, stream>>>: %r" % (cfg,)
        args_start = cfg_end + 3
        while text[args_start].isspace():
            args_start += 1
        args_end = _matching(text, args_start, "(", ")")
        name, targs = m.group(1), m.group(2) or ""
        assert name in barriers, "launch of an unknown kernel %s" % name
        out += text[pos:m.start()]
        out += "emul::launch(%s, %s, %s, [&] { %s%s%s; })" % (
            cfg[0], cfg[1], "true" if barriers[name] else "false", name, targs, text[args_start:args_end])
        pos = args_end
        count += 1
    return out + text[pos:], count, barriers


def _absolute_includes(text):
    def fix(m):
        return '#include "%s"' % os.path.normpath(os.path.join(CSRC, m.group(1)))
    return re.sub(r'#include\s+"([^"]+)"', fix, text)


def build(workdir):
    """-> (ctypes library, {kernel name: has_barrier}, number of rewritten launches)"""
    workdir = str(workdir)
    info, launches = {}, 0
    for src in ("assembly.cu", "taper.cu"):
        text, count, barriers = rewrite_launches(open(os.path.join(CSRC, src)).read())
        info.update(barriers)
        launches += count
        with open(os.path.join(workdir, src.replace(".cu", "_emul.inc")), "w") as f:
            f.write(_absolute_includes(text))
    so = os.path.join(workdir, "libcocons_host_emul.so")
    subprocess.check_call([
        "g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-pthread",
        "-I" + HERE, "-I" + workdir,
        '-DASSEMBLY_INC="assembly_emul.inc"', '-DTAPER_INC="taper_emul.inc"',
        os.path.join(HERE, "driver.cpp"), "-o", so])
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
    return lib, info, launches
