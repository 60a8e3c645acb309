#!/usr/bin/env python
"""Per-source-line / per-opcode summary of one kernel of an .ncu-rep captured with `--set full --import-source on`
(the library is compiled with -lineinfo).  Runs here, no GPU needed:

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--particles N] [--top 40] [--dump lines.tsv] [--kernel-index 0]

Prints: executed warp-instructions, thread-instructions per particle, opcode mix (share of instructions / of stall samples),
stall-reason totals, shared-memory wavefronts, per-file totals, the top lines by instructions and by samples.
--dump writes every line (file, line, warp-instr, thread-instr, samples, smem wavefronts, source text) as TSV.
"""
import argparse
import collections
import csv
import io
import subprocess
import sys


def export(rep, kernel_index):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    # one block per (kernel, file); blocks start with "File Path"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--particles", type=float, default=0)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--dump", default="")
    ap.add_argument("--kernel", default="", help="substring of the kernel name to keep (default: first kernel)")
    a = ap.parse_args()
    text = export(a.rep, 0)
    rows = list(csv.reader(io.StringIO(text)))
    cur_file, cur_func, header = None, None, None
    lines = collections.OrderedDict()   # (file, line) -> dict
    opc = collections.Counter()
    opc_samples = collections.Counter()
    stalls = collections.Counter()
    keep_func = None
    cur_key = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Function Name":
            cur_func = r[1]
            if keep_func is None and (not a.kernel or a.kernel in cur_func):
                keep_func = cur_func
            continue
        if r[0] == "Line No":
            header = r
            continue
        if header is None or cur_func != keep_func:
            continue
        d = dict(zip(header[4:], r[4:]))

        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        if r[0] != "":   # a source line row
            cur_key = (cur_file.split("/")[-1], int(r[0]))
            e = lines.setdefault(cur_key, dict(text=r[1], inst=0.0, tinst=0.0, samples=0.0, wave=0.0, wave_ideal=0.0))
            e["inst"] += num("Instructions Executed")
            e["tinst"] += num("Thread Instructions Executed")
            e["samples"] += num("# Samples")
            e["wave"] += num("L1 Wavefronts Shared")
            e["wave_ideal"] += num("L1 Wavefronts Shared Ideal")
        else:            # a SASS row under the current line
            sass = r[3].strip()
            op = sass.split()[0] if sass else "?"
            if op.startswith("@"):
                op = sass.split()[1]
            op = op.split(".")[0]
            opc[op] += num("Instructions Executed")
            opc_samples[op] += num("# Samples")
            for k in header:
                if k.startswith("stall_") and "(Not Issued)" not in k:
                    stalls[k] += num(k)
    tot = sum(e["inst"] for e in lines.values())
    ttot = sum(e["tinst"] for e in lines.values())
    stot = sum(e["samples"] for e in lines.values())
    wtot = sum(e["wave"] for e in lines.values())
    print(f"kernel: {keep_func}")
    print(f"warp-instructions {tot:.0f}  thread-instructions {ttot:.0f}  stall samples {stot:.0f}  smem wavefronts {wtot:.0f}")
    if a.particles:
        print(f"thread-instructions per particle {ttot / a.particles:.0f}   warp-instr per particle x32 {tot * 32 / a.particles:.0f}")
    print("opcode mix (inst share / sample share):")
    for op, n in opc.most_common(28):
        print(f"  {op:10s} {100 * n / tot:6.2f}%  {100 * opc_samples[op] / max(stot, 1):6.2f}%" + (f"  {n * 32 / a.particles:7.1f} thr-inst/particle" if a.particles else ""))
    st = sum(stalls.values())
    print("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / max(st, 1):.1f}%" for k, v in stalls.most_common(10)))
    files = collections.Counter()
    fs = collections.Counter()
    for (f, _), e in lines.items():
        files[f] += e["inst"]
        fs[f] += e["samples"]
    print("per file:")
    for f, n in files.most_common():
        print(f"  {f:34s} {100 * n / tot:6.2f}% inst {100 * fs[f] / max(stot, 1):6.2f}% samples" + (f" {n * 32 / a.particles:7.1f} thr-inst/particle" if a.particles else ""))
    print(f"top {a.top} lines by instructions:")
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["inst"])[: a.top]:
        print(f"  {f}:{ln:<5d} {100 * e['inst'] / tot:5.2f}% inst {100 * e['samples'] / max(stot, 1):5.2f}% smp  {e['text'].strip()[:110]}")
    print(f"top {a.top // 2} lines by samples:")
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[: a.top // 2]:
        print(f"  {f}:{ln:<5d} {100 * e['inst'] / tot:5.2f}% inst {100 * e['samples'] / max(stot, 1):5.2f}% smp  {e['text'].strip()[:110]}")
    if a.dump:
        with open(a.dump, "w") as f:
            f.write("file\tline\twarp_inst\tthread_inst\tsamples\tsmem_wavefronts\tsmem_wavefronts_ideal\ttext\n")
            for (fn, ln), e in sorted(lines.items()):
                f.write(f"{fn}\t{ln}\t{e['inst']:.0f}\t{e['tinst']:.0f}\t{e['samples']:.0f}\t{e['wave']:.0f}\t{e['wave_ideal']:.0f}\t{e['text'].strip()}\n")


if __name__ == "__main__":
    sys.exit(main())
