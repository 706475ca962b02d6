#!/usr/bin/env python
"""Runs the integer-pipe probes of lib/libzkb200_probe.so on cuda:0 and prints lane-ops/s.
usage: python tools/probe_run.py [kind ...]   (default: all kinds)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {0: "alu (VIADD+LOP3)", 1: "imad", 2: "alu+imad", 3: "prmt", 4: "shf", 5: "add64 pairs", 6: "blake2b compress/s",
         20: "LOP3 2 regs", 21: "LOP3 3 regs", 22: "IADD3 3 regs", 23: "IADD3 2 regs", 24: "PRMT 2 regs",
         25: "LOP3(2r)+IMAD(3r)", 26: "LOP3 2r + IADD3 2r", 27: "SHF 2 regs", 28: "IMAD 3 regs",
         30: "IMAD.HI 2 regs", 31: "IMAD.WIDE", 32: "IMAD reg*uniform+reg", 33: "IMAD.HI reg*uniform",
         40: "IMAD.WIDE reg*reg+acc64", 41: "IMAD.WIDE reg*uniform+acc64", 42: "IMAD.WIDE other-acc*reg+acc64",
         43: "IMAD.WIDE products only"}


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "zk_stark_tutor_b200", "lib", "libzkb200_probe.so"))
    kinds = [int(a) for a in sys.argv[1:]] or sorted(NAMES)
    for k in kinds:
        r, ms = ctypes.c_double(0), ctypes.c_double(0)
        rc = lib.zkb_probe_int_pipe(0, k, ctypes.byref(r), ctypes.byref(ms))
        print("kind %2d %-24s rc=%d  %.3f T/s  (%.3f ms)" % (k, NAMES.get(k, "?"), rc, r.value / 1e12, ms.value))


def blake():
    """In-register BLAKE2b compressions/s for each compression variant x occupancy cap."""
    lib = ctypes.CDLL(os.path.join(ROOT, "zk_stark_tutor_b200", "lib", "libzkb200_probe.so"))
    ref = None
    for v in (-1, 0, 32, 1, 2, 17):
        row = []
        for minb in (1, 2, 3, 4):
            r, ms, cs = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_uint64(0)
            rc = lib.zkb_probe_blake(0, v, minb, ctypes.byref(r), ctypes.byref(ms), ctypes.byref(cs))
            if ref is None:
                ref = cs.value
            row.append("%6.2f%s" % (r.value / 1e9, "" if (rc == 0 and cs.value == ref) else "(BAD rc=%d)" % rc))
        print("blake variant %3d  Gcompress/s at minb 1..4: %s" % (v, "  ".join(row)))


def blakex():
    """Pipe-balanced BLAKE2b family (blake2b_compress_x<CFG>): Gcompress/s per CFG."""
    lib = ctypes.CDLL(os.path.join(ROOT, "zk_stark_tutor_b200", "lib", "libzkb200_probe.so"))
    ref = None
    for cfg in (0, 1, 2, 3, 4, 8, 16, 28, 33, 35):          # the ten configurations still compiled into the probe library (25 in profiles/r01_probe.txt)
        r, ms, cs = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_uint64(0)
        rc = lib.zkb_probe_blakex(0, cfg, ctypes.byref(r), ctypes.byref(ms), ctypes.byref(cs))
        if ref is None:
            ref = cs.value
        print("blakex cfg %2d (add2:%d add3:%d r63:%d r24:%d r16:%d hi-imad:%d)  %6.2f Gcompress/s %s" % (
            cfg, cfg & 1, (cfg >> 1) & 1, (cfg >> 2) & 1, (cfg >> 3) & 1, (cfg >> 4) & 1, (cfg >> 5) & 1, r.value / 1e9,
            "" if (rc == 0 and cs.value == ref) else "(BAD rc=%d checksum %x vs %x)" % (rc, cs.value, ref)))


def blakey():
    """BLAKE2b with 64-bit adds as one accumulating IMAD.WIDE (blake2b_compress_y<CFG>): Gcompress/s per CFG."""
    lib = ctypes.CDLL(os.path.join(ROOT, "zk_stark_tutor_b200", "lib", "libzkb200_probe.so"))
    ref = None
    for cfg in (0, 2, 4, 29, 24, 49, 74, 99):               # the eight still compiled (24 in profiles/r02_probe_blakey.txt)
        r, ms, cs = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_uint64(0)
        rc = lib.zkb_probe_blakey(0, cfg, ctypes.byref(r), ctypes.byref(ms), ctypes.byref(cs))
        if ref is None:
            ref = cs.value
        print("blakey cfg %2d (c_frac:%d a_frac:%d hi-imad:%d a_mode:%d)  %6.2f Gcompress/s %s" % (
            cfg, cfg % 5, (cfg // 5) % 5, (cfg // 25) % 2, (cfg // 50) % 2, r.value / 1e9,
            "" if (rc == 0 and cs.value == ref) else "(BAD rc=%d checksum %x vs %x)" % (rc, cs.value, ref)))


def mix():
    """ALU-pipe xor + IMAD.WIDE / IMAD per step: clk per step per warp per scheduler (do the wide multiplies overlap with the ALU pipe?)."""
    lib = ctypes.CDLL(os.path.join(ROOT, "zk_stark_tutor_b200", "lib", "libzkb200_probe.so"))
    forms = {0: "IMAD.WIDE product", 1: "IMAD.WIDE accumulate", 2: "IMAD"}
    # (form 0 is not listed: ptxas rewrites a product whose high word is unused into a 32-bit IMAD)
    for a, w, f in ((16, 0, 0), (0, 8, 1), (0, 8, 2), (16, 8, 1), (16, 8, 2), (16, 4, 1), (16, 2, 1), (8, 8, 1), (8, 8, 2), (8, 4, 1), (4, 8, 1), (12, 8, 1)):
        r = ctypes.c_double(0)
        rc = lib.zkb_probe_mix(0, a, w, f, ctypes.byref(r))
        print("mix %2d LOP3 + %d %-20s rc=%d  %6.2f clk/step   (alone: ALU %d, multiplies %d at 2 clk / %d at 4 clk)" % (a, w, forms[f], rc, r.value, 2 * a, 2 * w, 4 * w))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mix":
        mix()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "blakey":
        blakey()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "blakex":
        blakex()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "blake":
        blake()
        sys.exit(0)
    main()
