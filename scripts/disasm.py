"""Print the node program the flattener emits for a built-in scene (or any Shader): python scripts/disasm.py npt-flange"""
import re
import struct
import sys

import numpy as np

sys.path.insert(0, ".")


def opcode_names():
    src = open("include/gsdf_program.h").read()
    body = src[src.index("enum gsdf_opcode {"):src.index("GSDF_OP__COUNT")]
    return [m for m in re.findall(r"^\s*(GSDF_OP_[A-Z0-9_]+)", body, re.M)]


def disasm(blob, file=sys.stdout):
    names = opcode_names()
    hdr = struct.unpack_from("<8I", blob, 0)
    words = np.frombuffer(blob, dtype=np.uint32, offset=32).reshape(-1, 4)
    print(f"nchunks={hdr[2]} dim={hdr[3]} dstack={hdr[4]} pstack={hdr[5]} ninstr={hdr[6]}", file=file)
    pc = 0
    out = []
    while pc < len(words):
        w = words[pc]
        op, ln = int(w[0]) & 0xff, (int(w[0]) >> 8) & 0xff
        name = names[op][len("GSDF_OP_"):]
        extra = ""
        if name in ("EXTRUDE_ENTER", "SCREW_ENTER") and int(w[1]) & 0xff:
            extra = f"  guard={int(w[1]) & 0xff} -> {int(w[1]) >> 8}"
        out.append((pc, name, ln, extra))
        print(f"{pc:4d}  {name:<18s} len={ln}{extra}", file=file)
        pc += ln
    return out


if __name__ == "__main__":
    from gsdf_b200 import gsdf as g
    scene = sys.argv[1] if len(sys.argv) > 1 else "npt-flange"
    bld = g.Builder()
    s = g.scene(bld, scene, float(sys.argv[2]) if len(sys.argv) > 2 else 0.0)
    disasm(bld.flatten(s)["blob"])
