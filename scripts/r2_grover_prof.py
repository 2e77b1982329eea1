"""A few Grover iterations for an ncu launch list (development aid): r2_grover_prof.py [qubits] [semantics]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcs_b200 import Circuit
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sem = sys.argv[2] if len(sys.argv) > 2 else "corrected"
sol = 0xABCDE % (1 << n)
c = Circuit(n, semantics=sem)
for q in range(n):
    c.h(q)
for _ in range(40):
    c.phase_flip(sol); c.diffusion()
c.flush()
print(c.get_probability(sol))
c.close()
