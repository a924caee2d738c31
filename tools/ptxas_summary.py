#!/usr/bin/env python
"""Summarise `-Xptxas -v` logs: kernel, registers, stack, spills, smem."""
import re, sys, glob
for path in sorted(sys.argv[1:] or glob.glob("build/obj/ptxas_*.log")):
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for '(\S+)'\s*\n(?:.*\n)*?.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\nptxas info\s*: Used (\d+) registers(.*)", txt):
        name = re.sub(r"^_ZN3vrb\d+", "", m.group(1)).replace("EEEvNS_9LaunchDevE", "")
        print(f"{name:50s} regs={m.group(6):>3s} stack={m.group(3):>3s} spill={m.group(4)}/{m.group(5)} {m.group(7).strip()}")
