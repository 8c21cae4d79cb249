#!/usr/bin/env python3
"""Turn one GPU round's scratch output (gpurun_out/*_<tag>.*) into the committed evidence under profiles/:

    python scripts/make_profiles.py r01f

  profiles/<tag>_ncu_<name>.txt      key metrics of each `ncu --set full` capture (from the .ncu-rep, raw page)
  profiles/<tag>_launches_<w>.csv    the ncu launch list (gpu__time_duration per launch) of bench.py --workload <w>
  profiles/<tag>_launch_shares.md    per-kernel share of the step from those launch lists
  profiles/<tag>_bench_<w>.json      the bench.py JSON lines of the round
  profiles/traffic_<w>.json          dram__bytes_read+write per launch of the dominant kernel (bench.py's `traffic`)
"""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
sys.path.insert(0, str(ROOT / "scripts"))
from ncu_summary import WANT  # noqa: E402

DOMINANT = {"cfg2": "k_fir_fast", "cfg3": "k_fir_fast", "cfg1": "k_demod_d", "chan": "k_chan_"}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main(tag):
    PROF.mkdir(exist_ok=True)
    for rep in sorted(OUT.glob(f"prof_*_{tag}.ncu-rep")):
        name = rep.stem[len("prof_"):-len(tag) - 1]
        hdr, units, rows = raw_rows(rep)
        lines = [f"# ncu --set full --clock-control none, capture {rep.name} (tag {tag})"]
        for vals in rows:
            lines.append(f"# kernel: {vals[hdr.index('Kernel Name')]}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append(f"{w:92s} {vals[i]:>18s} {units[i]}")
            w = next((t for t in name.split("_") if t in DOMINANT), None)
            if w in DOMINANT and DOMINANT[w] in vals[hdr.index("Kernel Name")]:
                def num(m):
                    i = hdr.index(m)
                    v = float(vals[i].replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
                tr = {"dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                      "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                      "kernel": vals[hdr.index("Kernel Name")], "capture": rep.name,
                      "note": "one ncu --set full capture at bench.py's full per-launch size"}
                (PROF / f"traffic_{w}.json").write_text(json.dumps(tr, indent=1) + "\n")
        (PROF / f"{tag}_ncu_{name}.txt").write_text("\n".join(lines) + "\n")
    shares = [f"# Per-kernel share of the step (ncu --metrics gpu__time_duration.sum launch lists, tag {tag})",
              "# ncu serialises launches and runs cold-cache: compare SHARES, not absolutes.", ""]
    for f in sorted(OUT.glob(f"launches_*_{tag}.csv")):
        w = f.stem[len("launches_"):-len(tag) - 1]
        shutil.copyfile(f, PROF / f"{tag}_launches_{w}.csv")
        rows = [r for r in csv.reader(open(f)) if len(r) > 10]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.defaultdict(list)
        for r in rows[1:]:
            try:
                agg[r[ki]].append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
        agg = {k: v for k, v in agg.items() if "synth_fill" not in k and "fold_taps" not in k}   # setup, not the step
        tot = sum(sum(v) for v in agg.values())
        shares.append(f"## {w}")
        shares.append("| kernel | launches | mean us | share of step |")
        shares.append("|---|---:|---:|---:|")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            shares.append(f"| `{k[:80]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |")
        shares.append("")
    (PROF / f"{tag}_launch_shares.md").write_text("\n".join(shares) + "\n")
    for f in sorted(OUT.glob(f"bench_*_{tag}*.json")):
        txt = f.read_text().strip().splitlines()
        if txt and txt[-1].startswith("{"):
            (PROF / f"{tag}_{f.stem[:-len(tag) - 1] if f.stem.endswith(tag) else f.stem}.json").write_text(txt[-1] + "\n")
    for f in list(OUT.glob(f"pytest_*_{tag}*.log")) + list(OUT.glob(f"smoke_{tag}.log")) + list(OUT.glob(f"gpu_{tag}.txt")):
        shutil.copyfile(f, PROF / f"{tag}_{f.name.replace('_' + tag, '')}")
    print("wrote", len(list(PROF.glob(f"{tag}_*"))), "files under profiles/")


if __name__ == "__main__":
    main(sys.argv[1])
