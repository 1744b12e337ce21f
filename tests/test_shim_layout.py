"""the Rust shim of INTEGRATION.md cannot be compiled here (no rustc), so its `#[repr(C)]`
structs and `extern "C"` declarations are checked as text: every struct must have the field order, offsets and size
of include/rtb200.h (measured with gcc: offsetof / sizeof) and of the ctypes mirror in _abi.py, and every declared
function must be exported by the library with the header's number of parameters."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALARS = {"u8": (1, 1), "u32": (4, 4), "i32": (4, 4), "u64": (8, 8), "i64": (8, 8), "f32": (4, 4), "f64": (8, 8),
           "c_int": (4, 4)}


def rust_structs():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\][^\n]*\n\s*pub struct (\w+)\s*\{(.*?)\n?\}", text, flags=re.S):
        body = re.sub(r"/\*.*?\*/", "", m.group(2), flags=re.S)
        fields = re.findall(r"pub (\w+)\s*:\s*([^,]+?)\s*(?:,|$)", body.strip(), flags=re.S)
        out[m.group(1)] = [(n, t.strip()) for n, t in fields]
    return out


def rust_layout(fields):
    """repr(C): each field at the next multiple of its alignment; size rounded up to the largest alignment."""
    off, max_al, res = 0, 1, []
    for name, ty in fields:
        arr = re.match(r"\[(\w+);\s*(\d+)\]", ty)
        if ty.startswith("*const") or ty.startswith("*mut"):
            size, al = 8, 8
        elif arr:
            s, al = SCALARS[arr.group(1)]
            size = s * int(arr.group(2))
        else:
            size, al = SCALARS[ty]
        off = (off + al - 1) // al * al
        res.append((name, off, size))
        off += size
        max_al = max(max_al, al)
    return res, (off + max_al - 1) // max_al * max_al


def c_layout(structs):
    """offsetof / sizeof of the same fields in include/rtb200.h, asked of the C compiler."""
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rtb200.h"', "int main(void) {"]
    for s, fields in structs.items():
        for n, _ in fields:
            lines.append('printf("%s %s %%zu %%zu\\n", offsetof(%s, %s), sizeof(((%s *)0)->%s));' % (s, n, s, n, s, n))
        lines.append('printf("%s . %%zu 0\\n", sizeof(%s));' % (s, s))
    lines += ["return 0;", "}"]
    src = "/tmp/rtb200_layout_probe.c"
    open(src, "w").write("\n".join(lines))
    exe = "/tmp/rtb200_layout_probe"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, src])
    out = {}
    for line in subprocess.check_output([exe], text=True).splitlines():
        s, n, a, b = line.split()
        out.setdefault(s, []).append((n, int(a), int(b)))
    return out


def test_repr_c_structs_match_header_and_ctypes(rt):
    structs = rust_structs()
    assert {"RtNode", "RtMaterial", "RtTexture", "RtPerlin", "RtImage", "RtSceneDesc", "RtCamera", "RtRenderOpts",
            "RtStats"} <= set(structs), sorted(structs)
    from_c = c_layout(structs)
    for name, fields in structs.items():
        lay, size = rust_layout(fields)
        c_fields, c_size = from_c[name][:-1], from_c[name][-1][1]
        assert lay == c_fields, (name, lay, c_fields)
        assert size == c_size, (name, size, c_size)
        ct = getattr(rt._abi, name)
        assert [f[0] for f in ct._fields_] == [n for n, _ in fields], name
        assert [(n, getattr(ct, n).offset, getattr(ct, n).size) for n, _ in fields] == lay, name
        assert C.sizeof(ct) == size, name


def _params(arglist):
    arglist = re.sub(r"/\*.*?\*/", "", arglist, flags=re.S).strip()
    if arglist in ("", "void"):
        return 0
    return arglist.count(",") + 1


def test_extern_c_declarations_match_header_and_library(rt):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    block = re.sub(r"//[^\n]*", "", block)
    rust = {m.group(1): _params(m.group(2)) for m in re.finditer(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->[^;]+)?;", block, flags=re.S)}
    assert {"rt_scene_create", "rt_render", "rt_scene_destroy", "rt_last_error"} <= set(rust), sorted(rust)
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rtb200.h")).read(), flags=re.S)
    proto = {m.group(1): _params(m.group(2)) for m in re.finditer(r"\b(rt_\w+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)}
    for name, n in rust.items():
        assert name in proto, "%s is not declared in include/rtb200.h" % name
        assert proto[name] == n, (name, n, proto[name])
        assert hasattr(rt._dev, name), "%s is not exported by librtb200.so" % name
