#!/usr/bin/env python
"""Turn an .ncu-rep (from `ncu --set full --import-source on`, brought back in gpurun_out/) into the
small text summary that is committed under profiles/:  per-kernel headline metrics (duration, DRAM
bytes, throughput percentages, occupancy limiters), the executed-instruction histogram by SASS
opcode, and the CUDA source lines that execute the most instructions / collect the most stall samples.

    python profiles/summarize.py gpurun_out/prof_r1d.ncu-rep > profiles/r01_micro2048_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, check=True, capture_output=True, text=True).stdout


def raw_page(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        out.append(d)
    return out


def source_pages(rep):
    """-> {function name: {"ops": Counter, "total": int, "lines": [(insts, samples, file:line, text)]}}
    parsed from the cuda,sass view: per file, a CUDA line row (Line No set) is followed by its SASS rows"""
    txt = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    out, fpath, fn, hdr = {}, None, None, None
    nth = collections.Counter()            # a function profiled several times repeats its files: keep the first launch
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            fn = r[1]
            nth[(fn, fpath)] += 1
            out.setdefault(fn, {"ops": collections.Counter(), "total": 0, "lines": []})
        elif r[0] == "Line No":
            hdr = r
            i_ex, i_smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif hdr and len(r) >= len(hdr) - 1 and fn:
            try:
                n, smp = int(r[i_ex]), int(r[i_smp])
            except ValueError:
                continue
            if r[0] and nth[(fn, fpath)] == 1:      # a CUDA source line (aggregate of its SASS)
                out[fn]["lines"].append((n, smp, "%s:%s" % (fpath, r[0]), r[1].strip()))
    # opcode histogram from the plain sass view (every instruction exactly once)
    txt = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"])
    fn, hdr = None, None
    seen = collections.Counter()
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            fn = r[1]
            seen[fn] += 1
            hdr = None
        elif r[0] == "Address":
            hdr = r
            i_ex = hdr.index("Instructions Executed")
        elif hdr and fn in out and seen[fn] == 1 and r[0].startswith("0x"):
            try:
                n = int(r[i_ex])
            except ValueError:
                continue
            s = r[1].split()
            op = s[1] if s[0].startswith("@") else s[0]
            out[fn]["ops"][op.split(".")[0]] += n
            out[fn]["total"] += n
    return out


def main():
    rep = sys.argv[1]
    raws = raw_page(rep)
    srcs = source_pages(rep)
    done = set()
    print("# ncu summary of %s (ncu --set full --clock-control none --import-source on)" % rep.split("/")[-1])
    print("# per-launch numbers under ncu are cold-cache and serialised: use shares and counts, not absolute speed\n")
    for i, d in enumerate(raws):
        print("=" * 100)
        print("launch %d: %s" % (i, d["Kernel Name"][0]))
        for m in RAW:
            if m in d:
                print("  %-72s %16s %s" % (m, d[m][0], d[m][1]))
        key = [k for k in srcs if k.replace("dbde::", "").split("(")[0].split("<")[0].split()[-1] ==
               d["Kernel Name"][0].split("(")[0].split("<")[0].split()[-1]]
        if key and key[0] not in done:
            done.add(key[0])
            S = srcs[key[0]]
            if S["total"]:
                print("  -- executed warp instructions by opcode (total %d; first launch of this kernel)" % S["total"])
                for op, n in S["ops"].most_common(18):
                    print("     %-12s %12d  %5.1f%%" % (op, n, 100.0 * n / S["total"]))
            if S["lines"]:
                print("  -- CUDA lines by executed warp instructions (top 45): insts, stall samples, file:line, source")
                for n, smp, where, text in sorted(S["lines"], reverse=True)[:45]:
                    print("     %11d %7d  %-24s %s" % (n, smp, where, text[:100]))
        print()


if __name__ == "__main__":
    main()
