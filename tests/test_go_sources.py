"""The Go side of the drop-in (integration/go/) cannot be compiled in this image (no Go toolchain), so these checks keep
it consistent with what CAN be checked here: the opcode and guard enums of include/gsdf_program.h, the C ABI of
include/gsdf_b200.h that the cgo files call, the built library's exports, and -- when the reference tree is present --
the list of node types that need an emitter and the struct fields the emitters read."""
import glob
import os
import re

import pytest

from gsdf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "integration", "go")
REF = "/root/reference"


def read(*parts):
    with open(os.path.join(*parts)) as f:
        return f.read()


def strip_go(src):
    """Go source without comments, string and rune literals (enough for bracket counting)."""
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r'"(\\.|[^"\\\n])*"', '""', src)
    src = re.sub(r"`[^`]*`", '""', src)
    src = re.sub(r"'(\\.|[^'\\\n])+'", "' '", src)
    return src


def test_go_opcodes_follow_the_c_enum():
    hdr = read(ROOT, "include", "gsdf_program.h")
    enum = hdr[hdr.index("enum gsdf_opcode"):]
    enum = enum[:enum.index("};")]
    c_ops = re.findall(r"\b(GSDF_OP_[A-Z0-9_]+)\b\s*(?:=\s*0\s*)?,", enum)
    c_ops = [o for o in c_ops if o != "GSDF_OP__COUNT"]
    go = strip_go(read(GO, "glbuild", "cuda_program.go"))
    block = go[go.index("OpEnd uint32 = iota"):]
    block = block[:block.index(")")]
    go_ops = re.findall(r"^\s*(Op\w+)", block, flags=re.M)
    assert go_ops[0] == "OpEnd" and len(go_ops) >= 50
    norm = lambda s: s.replace("GSDF_OP_", "").replace("_", "").lower()
    # the Go list IS the C enum (box guards and the fold seed included: the Go flattener emits them like the C++ one)
    assert [g[2:].lower() for g in go_ops] == [norm(c) for c in c_ops]
    # guard kinds, magic, version
    kinds = re.search(r"enum gsdf_guard_kind \{([^}]*)\}", hdr).group(1)
    kinds = [k.split("=")[0].strip() for k in kinds.split(",")]
    gblock = go[go.index("GuardNone uint32 = iota"):]
    gblock = re.findall(r"^\s*(Guard\w+)", gblock[:gblock.index(")")], flags=re.M)
    assert [g[5:].lower() for g in gblock] == [k.replace("GSDF_GUARD_", "").replace("_", "").lower() for k in kinds]
    assert re.search(r"define GSDF_PROGRAM_MAGIC (0x[0-9a-fA-F]+)", hdr).group(1).lower() == re.search(r"programMagic\s*=\s*(0x[0-9a-fA-F]+)", go).group(1).lower()
    assert re.search(r"define GSDF_PROGRAM_VERSION (\d+)", hdr).group(1) == re.search(r"programVersion\s*=\s*(\d+)", go).group(1)


def test_cgo_files_only_call_declared_and_exported_symbols():
    hdr = read(ROOT, "include", "gsdf_b200.h")
    declared = set(re.findall(r"\b(gsdf_\w+)\s*\(", hdr)) | set(re.findall(r"\}\s*(gsdf_\w+)\s*;", hdr)) | \
        set(re.findall(r"typedef struct (gsdf_\w+)", hdr))
    used = set()
    for path in glob.glob(os.path.join(GO, "*", "*.go")):
        used |= set(re.findall(r"\bC\.(gsdf_\w+)", read(path)))
    assert used and used <= declared, sorted(used - declared)
    import ctypes
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for fn in used:
        if re.search(r"\b%s\s*\(" % fn, hdr):      # a function, not a type
            assert hasattr(lib, fn), fn


def test_go_files_are_bracket_balanced_and_packaged():
    want_pkg = {"glbuild": "glbuild", "gsdf": "gsdf", "gleval": "gleval", "glrender": "glrender", "threads": "threads"}
    files = glob.glob(os.path.join(GO, "*", "*.go")) + glob.glob(os.path.join(GO, "*", "*", "*.go"))
    assert len(files) >= 6
    for path in files:
        src = read(path)
        pkg = re.search(r"^package (\w+)", src, flags=re.M).group(1)
        assert pkg == want_pkg[os.path.basename(os.path.dirname(path))], path
        body = strip_go(src)
        for a, b in ("{}", "()", "[]"):
            assert body.count(a) == body.count(b), (path, a, body.count(a), body.count(b))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_every_reference_node_type_has_an_emitter_reading_real_fields():
    ref = ""
    for f in glob.glob(os.path.join(REF, "*.go")) + glob.glob(os.path.join(REF, "forge", "threads", "*.go")):
        if not f.endswith("_test.go"):
            ref += read(f) + "\n"
    # node types = receivers of a CPU Evaluate method (cpu_evaluators.go, forge/threads/threads.go)
    nodes = set(re.findall(r"func \(\w+ \*(\w+)\) Evaluate\(", read(REF, "cpu_evaluators.go")))
    nodes |= set(re.findall(r"func \(\w+ \*(\w+)\) Evaluate\(", read(REF, "forge", "threads", "threads.go")))
    go = read(GO, "gsdf", "cuda_flatten.go") + read(GO, "forge", "threads", "cuda_flatten.go")
    emitters = set(re.findall(r"func \(\w+ \*(\w+)\) AppendProgram\(", go))
    assert nodes <= emitters, sorted(nodes - emitters)
    # every receiver field an emitter reads exists on the reference struct (embedded structs included)
    structs = {}
    for m in re.finditer(r"type (\w+) struct \{(.*?)\n\}", ref, flags=re.S):
        fields = set()
        for line in m.group(2).split("\n"):
            parts = line.split("//")[0].split()
            if len(parts) == 1:
                fields.add("EMBED:" + parts[0].lstrip("*").split(".")[-1])
            elif parts:
                fields |= {n.strip() for n in " ".join(parts[:-1]).split(",") if n.strip()}
        structs[m.group(1)] = fields

    def all_fields(t):
        fs = set(structs.get(t, ()))
        for f in list(fs):
            if f.startswith("EMBED:"):
                fs |= all_fields(f[6:])
        return fs
    methods = set(re.findall(r"func \(\w+ \*?(\w+)\) (\w+)\(", ref + go))
    for chunk in re.split(r"\n(?=func )", go):
        m = re.match(r"func \((\w+) \*(\w+)\) \w+\(", chunk)
        if not m or m.group(1) == "_":
            continue
        recv, typ = m.groups()
        assert typ in structs, typ
        for used in set(re.findall(r"\b%s\.(\w+)" % recv, chunk)):
            assert used in all_fields(typ) or (typ, used) in methods, (typ, used)


def test_go_emitters_use_the_opcodes_the_cpp_flattener_uses_per_node_type():
    """Every node type: the opcodes its Go AppendProgram emits are the opcodes its case in flatten.cpp emits (helpers that
    exist on both sides -- position push / pop, box-guard emission -- hide the same ops on both sides). Catches an emitter
    that forgets an exit op, a guard or the fold seed, or uses another combiner than the executable specification."""
    go = read(GO, "gsdf", "cuda_flatten.go") + read(GO, "forge", "threads", "cuda_flatten.go")
    cpp = read(ROOT, "gsdf_b200", "csrc", "host", "flatten.cpp")
    circ = re.findall(r"glbuild\.Op(\w+)", go[go.index("func circ("):go.index("func (u *circarray) AppendProgram")])
    gops = {}
    for chunk in re.split(r"\n(?=func )", go):
        m = re.match(r"func \((\w+) \*(\w+)\) AppendProgram\(", chunk)
        if not m:
            continue
        ops = re.findall(r"glbuild\.Op(\w+)", chunk) + (circ if "circ(p" in chunk else [])
        gops[m.group(2).lower()] = {o.lower() for o in ops}
    body = cpp[cpp.index("bool emit(NodeId id"):cpp.index("// Radius reuse")]
    cases = list(re.finditer(r"((?:\s*case GSDF_N_\w+:)+)", body))
    cops = {}
    for i, m in enumerate(cases):
        blk = body[m.end(): cases[i + 1].start() if i + 1 < len(cases) else len(body)]
        ops = {o.lower().replace("_", "") for o in re.findall(r"GSDF_OP_(\w+)", blk)}
        for k in re.findall(r"GSDF_N_(\w+)", m.group(1)):
            cops[k.lower().replace("_", "")] = ops
    alias = {"opunion": "union", "opunion2d": "union2d", "diamond": "diamond2d", "equilateraltri2d": "eqtri2d", "extrusion": "extrude",
             "quadbezier2d": "bezierq2d", "revolution": "revolve", "rotation2d": "rotate2d", "x2d": "roundx2d"}
    checked = 0
    for typ, ops in gops.items():
        kind = alias.get(typ, typ)
        assert kind in cops, typ
        want = set(cops[kind])
        if kind in ("array", "array2d"):          # one C++ case serves both node kinds
            want -= {"array2dvar"} if kind == "array" else {"arrayvar"}
        if kind in ("union", "union2d"):          # likewise: the 2-D box-guard ops sit in the shared union case
            want -= {"cullub2d"} if kind == "union" else set()
        assert ops == want, (typ, sorted(ops ^ want))
        checked += 1
    assert checked >= 50


def test_cgo_calls_pass_as_many_arguments_as_the_prototypes_take():
    hdr = re.sub(r"/\*.*?\*/", "", read(ROOT, "include", "gsdf_b200.h"), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(gsdf_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    assert len(protos) >= 40

    def top_level_args(s):
        depth, cur, out = 0, "", []
        for ch in s:
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 0:
                out.append(cur)
                cur = ""
            else:
                cur += ch
        return out + ([cur] if cur.strip() else [])
    calls = 0
    for path in glob.glob(os.path.join(GO, "*", "*.go")):
        src = strip_go(read(path))
        for m in re.finditer(r"C\.(gsdf_\w+)\(", src):
            if m.group(1) not in protos:
                continue            # a type conversion such as C.gsdf_lattice(...)
            i = j = m.end()
            depth = 1
            while depth:
                depth += src[j] == "("
                depth -= src[j] == ")"
                j += 1
            assert len(top_level_args(src[i:j - 1])) == protos[m.group(1)], (path, m.group(1))
            calls += 1
    assert calls >= 15
