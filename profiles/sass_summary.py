#!/usr/bin/env python
"""Blackwell-native SASS evidence: counts of tcgen05 / TMA / bulk-copy / TMEM / REDUX mnemonics per kernel in librtkv.so
(`cuobjdump -sass`).  usage: python profiles/sass_summary.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "realtime-kv-cache-compression_b200", "rtkv", "librtkv.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTMALDG\.\dD|UTMASTG|UBLKCP\.S\.G|LDTM\.x\d+|STTM\.x\d+|UTCBAR|SYNCS\.[A-Z0-9.]+|REDUX[.\w]*|CREDUX\.\w+|"
                 r"MEMBAR\.[A-Z.]+|HADD2\.BF16_V2|HMUL2\.BF16_V2|HMNMX2\.BF16_V2[.\w]*|HFMA2\.BF16_V2|MUFU\.EX2)\b")
KERNELS = ("flash_prefill", "prompt_mass_qk", "shard_stats_staged_kernel", "shard_combine", "quant_gather_flat", "quant_gather_vec_kernelILi2ELi2",
           "quant_gather_vec_kernelILi0ELi4", "score_select_cluster", "prompt_mass_kernelILi0ELb1")
txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if not any(k in name for k in KERNELS):
        continue
    c = collections.Counter(m.group(0) for m in PAT.finditer(f))
    print(name)
    print("    " + ", ".join(f"{k} x{v}" for k, v in c.most_common()))
