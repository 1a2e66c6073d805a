#!/usr/bin/env python
"""Derives an egg whose morphs reach beyond positions and normals from the reference's models/ripple.egg (which only
has <Dxyz> / <DNormal>): every vertex gets an <RGBA> with <DRGBA> deltas and its <UV> gets <Duv> deltas for the
sliders the vertex already has.  The egg loader then emits a colour column (4 x uint8, or packed_dabc with
`vertex-colors-prefer-packed 1`) and a texcoord column as MORPH BASES (egg2pg/eggLoader.cxx:2392-2412) -- the
loader-made formats whose animation goes through the colour packers (gobj/geomVertexColumn.cxx:4465-4552).

    python integration/make_morph_egg.py <ripple.egg> <out.egg>

Written into integration/_build/models/ by integration/build.sh; nothing of the reference is committed.
"""
import re
import sys


def main(src, dst):
    out = []
    rng = 12345
    def rnd():   # small deterministic LCG, values in [-0.5, 0.5)
        nonlocal rng
        rng = (rng * 1103515245 + 12345) & 0x7FFFFFFF
        return rng / 0x7FFFFFFF - 0.5
    lines = open(src).read().split("\n")
    i = 0
    while i < len(lines):
        line = lines[i]
        if re.match(r"\s*<Vertex> \d+ \{", line):
            # collect the vertex block
            depth, block = 0, []
            while True:
                block.append(lines[i])
                depth += lines[i].count("{") - lines[i].count("}")
                i += 1
                if depth == 0:
                    break
            sliders = sorted({m.group(1) for b in block for m in [re.search(r"<Dxyz> (\S+) \{", b)] if m})
            new_block = []
            for b in block:
                m = re.match(r"(\s*)<UV> \{ (\S+) (\S+) \}", b)
                if m and sliders:
                    ind = m.group(1)
                    new_block.append(f"{ind}<UV> {{")
                    new_block.append(f"{ind}  {m.group(2)} {m.group(3)}")
                    for s in sliders:
                        new_block.append(f"{ind}  <Duv> {s} {{ {rnd() * 0.2:.6f} {rnd() * 0.2:.6f} }}")
                    new_block.append(f"{ind}}}")
                else:
                    new_block.append(b)
            ind = "    "
            rgba = [f"{ind}<RGBA> {{", f"{ind}  {0.5 + rnd() * 0.8:.6f} {0.5 + rnd() * 0.8:.6f} {0.5 + rnd() * 0.8:.6f} {0.75 + rnd() * 0.4:.6f}"]
            for s in sliders:
                rgba.append(f"{ind}  <DRGBA> {s} {{ {rnd():.6f} {rnd():.6f} {rnd():.6f} {rnd() * 0.5:.6f} }}")
            rgba.append(f"{ind}}}")
            new_block[-1:-1] = rgba          # before the closing brace of the vertex
            out.extend(new_block)
            continue
        out.append(line)
        i += 1
    open(dst, "w").write("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
