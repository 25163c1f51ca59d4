"""Prints the pipe microbenchmarks of this GPU (IMAD, IMAD.WIDE, FP64 FMA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shielded_pool_pinocchio_solana_b200 as g16
ctx = g16.Context(0)
for kind, name in ((0, "IMAD"), (1, "IMAD.WIDE"), (2, "DFMA")):
    print(name, "%.2f T instr/s" % (ctx.measure_imad_peak(kind) / 1e12), flush=True)
