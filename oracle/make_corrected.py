#!/usr/bin/env python3
"""oracle/make_corrected.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Writes a *patched scratch copy* of the reference's src/ tree into a directory
OUTSIDE this repository (default /tmp/qcs_ref_corrected/src) with the three
arithmetic defects of SURVEY.md section 0.2 repaired:

  D1  q_apply_2q_gate: the `else { scratch[i] = vector[i]; }` arm overwrites the
      freshly computed scratch[i | t_bit] (src/q_gates.c:289-291 and the four
      other loop variants + the pthread worker).  Fix: drop the arm -- scratch
      already holds a full copy of vector from the memcpy at the loop head.
  D2  q_apply_phase_flip swaps vector/scratch before negating one element
      (src/q_gates.c:311-316), exposing the stale scratch buffer.  Fix: negate
      vector[index] in place, no swap.
  D3  q_apply_diffusion divides the amplitude sum by its own magnitude instead
      of by N (src/q_gates.c:337-344).  Fix: mean = sum / (double)size.

The edits are regular-expression rewrites applied at build time; no reference
source text is stored in this repository.  oracle/Makefile compiles the result
into oracle/_ref/libqcsref_corrected.so ("corrected" semantics oracle).
"""
import os
import re
import sys


def patch_q_gates(text: str) -> str:
    # D1: remove every else-arm that copies vector[i] into scratch[i].
    pat_d1 = re.compile(
        r"\}\s*else\s*\{\s*state->scratch_vector\[i\]\s*=\s*state->vector\[i\];\s*\}"
    )
    text, n1 = pat_d1.subn("}", text)
    if n1 < 6:
        raise SystemExit(f"make_corrected: expected >=6 D1 sites, patched {n1}")

    # D2: phase flip in place (drop the pointer swap, negate vector[index]).
    pat_d2 = re.compile(
        r"(void q_apply_phase_flip\(struct t_q_state \*state, int index\) \{.*?return;\s*\})"
        r"(.*?)\n\}",
        re.S,
    )
    m = pat_d2.search(text)
    if not m:
        raise SystemExit("make_corrected: q_apply_phase_flip not found")
    body = (
        "\n\n  state->vector[index].number_real = -state->vector[index].number_real;\n"
        "  state->vector[index].number_imaginary = "
        "-state->vector[index].number_imaginary;"
    )
    text = text[: m.start(2)] + body + text[m.end(2):]

    # D3: mean = sum / N  (replace the magnitude normalisation block).
    pat_d3 = re.compile(
        r"if \(mean_magnitude > 1e-10\) \{\s*mean\.number_real /= mean_magnitude;\s*"
        r"mean\.number_imaginary /= mean_magnitude;\s*\}"
    )
    text, n3 = pat_d3.subn(
        "mean.number_real /= (double)size;\n  mean.number_imaginary /= (double)size;",
        text,
    )
    if n3 != 1:
        raise SystemExit(f"make_corrected: expected 1 D3 site, patched {n3}")
    return text


def main() -> None:
    ref_src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src"
    out_dir = sys.argv[2] if len(sys.argv) > 2 else "/tmp/qcs_ref_corrected/src"
    here = os.path.realpath(os.path.dirname(__file__))
    repo = os.path.dirname(here)
    if os.path.realpath(out_dir).startswith(repo + os.sep):
        raise SystemExit("make_corrected: refusing to write reference text into the repo")
    os.makedirs(out_dir, exist_ok=True)
    inc_dir = os.path.join(os.path.dirname(out_dir), "include")
    os.makedirs(inc_dir, exist_ok=True)
    for name in sorted(os.listdir(ref_src)):
        if not (name.endswith(".c") or name.endswith(".h")):
            continue
        with open(os.path.join(ref_src, name), "r", encoding="utf-8") as f:
            text = f.read()
        if name == "q_gates.c":
            text = patch_q_gates(text)
        with open(os.path.join(out_dir, name), "w", encoding="utf-8") as f:
            f.write(text)
    ref_inc = os.path.join(os.path.dirname(ref_src), "include")
    for name in os.listdir(ref_inc):
        with open(os.path.join(ref_inc, name), "r", encoding="utf-8") as f:
            text = f.read()
        with open(os.path.join(inc_dir, name), "w", encoding="utf-8") as f:
            f.write(text)
    print(out_dir)


if __name__ == "__main__":
    main()
