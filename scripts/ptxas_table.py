#!/usr/bin/env python
"""Registers / spill bytes per kernel from a `ptxas -v` log.  usage: ptxas_table.py <log> [substring]"""
import re
import sys

log = open(sys.argv[1]).read()
want = sys.argv[2] if len(sys.argv) > 2 else ""
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_\w+'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, "
                     r"(\d+) bytes spill loads.*?Used (\d+) registers", log, re.S):
    name, stack, st, ld, regs = m.groups()
    if want in name:
        short = re.sub(r"^_ZN3bgr", "", name)
        short = re.sub(r"EEvNS_11RolloutArgsE$", "", short)
        print(f"{short:60s} regs {regs:>3s}  stack {stack:>4s}  spill st/ld {st:>4s}/{ld:>4s}")
