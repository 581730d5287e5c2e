"""Prototype: emit the straight-line specialisation of a node program (one exec_one call per instruction with the opcode word
and the flag word as constants; operands stay in shared memory). python scripts/gen_special.py npt-flange > special.cuh"""
import struct, sys
import numpy as np
sys.path.insert(0, ".")
from gsdf_b200 import gsdf


def gen(blob, name="run_special"):
    hdr = struct.unpack_from("<8I", blob, 0)
    words = np.frombuffer(blob, dtype=np.uint32, offset=32).reshape(-1, 4)
    out = ["template <int P, bool EXT>", "__device__ __forceinline__ void %s(Machine<P> &m, const uint4 *__restrict__ prog, const float4 *__restrict__ aux) {" % name,
           "    int pc = 0;", "    uint4 h;"]
    pc = 0
    targets = set()
    ins = []
    while pc < len(words):
        w = words[pc]
        op, ln = int(w[0]) & 0xff, (int(w[0]) >> 8) & 0xff
        ins.append((pc, int(w[0]), int(w[1]), ln))
        pc += ln
    # guard targets: any instruction whose flag word carries a target in bits 8.. and a guard kind in bits 0..7 MAY jump;
    # emit a label for every instruction and let the compiler drop the unused ones
    for pc, w0, w1, ln in ins:
        out.append("L%d:" % pc)
        out.append("    h = prog[%d]; h.x = 0x%xu; h.y = 0x%xu; pc = %d;" % (pc, w0, w1, pc))
        if (w0 & 0xff) == 0:
            out.append("    return;")
            continue
        out.append("    exec_one<P, EXT>(m, h, prog, pc, aux);")
        tgt = w1 >> 8
        if (w1 & 0xff) and tgt > pc and tgt in [i[0] for i in ins]:
            out.append("    if (pc != %d) goto L%d;" % (pc + ln, tgt))
    out.append("}")
    return "\n".join(out)


if __name__ == "__main__":
    b = gsdf.Builder()
    s = gsdf.scene(b, sys.argv[1] if len(sys.argv) > 1 else "npt-flange")
    print(gen(b.flatten(s)["blob"]))
