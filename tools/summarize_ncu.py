"""Turn an .ncu-rep (read here, no GPU needed) into the text summary committed under profiles/."""
import collections, csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_tmem.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
want_sub = ["sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"]
with open(out, "w") as f:
    f.write(f"# summary of {rep.split('/')[-1]} (ncu --set full --clock-control none --import-source on)\n")
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                f.write(f"{w:72s} {units[i]:16s} {[r[i] for r in data]}\n")
    for w in want_sub:
        for i, h in enumerate(hdr):
            if h.endswith(w):
                f.write(f"{h:72s} {units[i]:16s} {[r[i] for r in data]}\n")
    rows = list(csv.reader(io.StringIO(src)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}; secs.append(cur); continue
        if cur is not None: cur["rows"].append(r)
    for sec in secs[:1]:
        h = sec["rows"][0]; ix = {k: i for i, k in enumerate(h)}
        d = [r for r in sec["rows"][1:] if len(r) >= len(h) and r[0].startswith("0x")]
        I = lambda r, k: int(r[ix[k]] or 0)
        stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        n = sum(I(r, "# Samples") for r in d)
        tot = collections.Counter()
        for r in d:
            for s in stalls: tot[s] += I(r, s)
        f.write(f"\n# warp-state samples ({n} total)\n")
        for k, v in tot.most_common(10): f.write(f"  {k:28s} {100 * v / n:5.1f} %\n")
        def opc(s):
            p = s.strip().split(); o = p[1] if p[0].startswith("@") else p[0]; return o.split(".")[0]
        opn = collections.Counter()
        for r in d: opn[opc(r[ix["Source"]])] += I(r, "Instructions Executed")
        ti = sum(opn.values())
        f.write("\n# executed warp instructions by opcode\n")
        for k, v in opn.most_common(10): f.write(f"  {k:10s} {v:14d}  {100 * v / ti:5.1f} %\n")
        wf = sum(I(r, "L1 Wavefronts Shared") for r in d); ex = sum(I(r, "L1 Wavefronts Shared Excessive") for r in d)
        f.write(f"\n# shared-memory wavefronts {wf}, excessive (bank conflicts) {ex} ({100 * ex / max(wf, 1):.1f} %)\n")
        tma = [r[ix["Source"]].strip() for r in d if any(t in r[ix["Source"]] for t in ("UTMALDG", "UBLKCP", "SYNCS"))]
        f.write("\n# TMA / bulk-copy / mbarrier instructions in the SASS\n")
        for t in tma[:12]: f.write("  " + t + "\n")
        tc = collections.Counter()
        for r in d:
            o = opc(r[ix["Source"]])
            if o.startswith("UTC") or o in ("LDTM", "STTM"): tc[r[ix["Source"]].strip().split("(")[0][:60].split(",")[0].split(" [")[0]] += I(r, "Instructions Executed")
        if tc:
            f.write("\n# tensor-core / tensor-memory instructions executed (tcgen05: UTC*MMA, UTCBAR = commit, LDTM = tcgen05.ld)\n")
            for k, v in tc.most_common(12): f.write(f"  {k:44s} {v:12d}\n")
print(open(out).read())
