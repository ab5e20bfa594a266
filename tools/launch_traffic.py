"""ncu launch-list CSV of ONE eager cfg2 train step -> profiles/<name>.md (time + DRAM bytes per kernel) and
profiles/traffic.json (DRAM bytes per launch, read by bench.py for `roofline.traffic`).

  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-steps 1 --no-graph
  python tools/launch_traffic.py gpurun_out/launches.csv profiles/r2_launches_final.md profiles/traffic.json

The attention kernels are split by launch shape: attn1 (6144 keys: 1-D grid over the (key tile, head) items, or
(q tile, head) for the forward) and attn2 (256 caption keys) run the same kernel symbols."""
import collections
import csv
import json
import re
import sys


def read(path):
    rows = collections.OrderedDict()
    lines = [l for l in open(path) if not l.startswith("==")]
    for r in csv.DictReader(lines):
        e = rows.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]), "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r["Metric Unit"], v)
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
            e[r["Metric Name"]] = v * mult
    return list(rows.values())


def family(e):
    """bench.py's kernel keys (ops.KernelTimer families, with the launch shape for the attention kernels)."""
    n = e["name"]
    big = e.get("us", 0.0) > 150.0      # attn1 at 6144 x 6144 runs 390-870 us, attn2 (256 keys) 25-40 us
    if "fa_bwd_kernel" in n:
        return "fa_bwd 1x32x6144x6144" if big else "fa_bwd_attn2 1x32x6144x256"
    if "fa_fwd_db_kernel" in n:
        return "fa_fwd 1x32x6144x6144"
    if "fa_fwd_kernel" in n:
        return "fa_fwd_attn2 1x32x6144x256"
    for k in ("gemm", "norm_mod_bwd", "norm_mod_fwd", "qknorm_rope_bwd", "qknorm_rope_fwd", "attn_delta", "rowscale"):
        if "b200::" + k in n:
            return k
    return None


def main(src, out_md, out_json, note=""):
    rows = read(src)
    agg = collections.defaultdict(lambda: {"us": [], "rd": 0.0, "wr": 0.0})
    fam = collections.defaultdict(lambda: {"n": 0, "bytes": 0.0})
    for e in rows:
        a = agg[e["name"][:80]]
        a["us"].append(e.get("us", 0.0))
        a["rd"] += e.get("dram__bytes_read.sum", 0.0)
        a["wr"] += e.get("dram__bytes_write.sum", 0.0)
        f = family(e)
        if f:
            fam[f]["n"] += 1
            fam[f]["bytes"] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
    tot = sum(sum(a["us"]) for a in agg.values())
    with open(out_md, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum + DRAM bytes, --clock-control none; cold-cache, serialised)\n\n")
        if note:
            f.write(note.rstrip() + "\n\n")
        f.write(f"total kernel time {tot / 1e3:.2f} ms over {sum(len(a['us']) for a in agg.values())} launches, "
                f"{sum(len(a['us']) for k, a in agg.items() if 'b200::' in k)} of them b200 kernels\n\n")
        f.write("| share | launches | avg us | min us | max us | DRAM read MB / launch | DRAM write MB / launch | kernel |\n"
                "|---|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -sum(kv[1]["us"])):
            v = a["us"]
            f.write(f"| {sum(v) / tot * 100:.1f}% | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {max(v):.1f} | "
                    f"{a['rd'] / len(v) / 1e6:.1f} | {a['wr'] / len(v) / 1e6:.1f} | `{k}` |\n")
        f.write("\nDRAM bytes per launch by bench.py kernel key (`profiles/traffic.json`):\n\n| key | launches | MB / launch |\n|---|---|---|\n")
        for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["bytes"]):
            f.write(f"| `{k}` | {d['n']} | {d['bytes'] / d['n'] / 1e6:.1f} |\n")
    tj = {k: d["bytes"] / d["n"] for k, d in fam.items()}
    tj["fa_bwd"] = tj.get("fa_bwd 1x32x6144x6144")
    tj["fa_fwd"] = tj.get("fa_fwd 1x32x6144x6144")
    tj["_source"] = (f"{out_md} ({src}; ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over ONE eager cfg2 train "
                     "step): read + write bytes averaged per launch; keys with a shape are single kernels at that launch "
                     "shape (the bench line's roofline.kernel), bare family keys average all launches of the family")
    json.dump(tj, open(out_json, "w"), indent=1)
    print(open(out_md).read()[:2500])


if __name__ == "__main__":
    main(*sys.argv[1:5])
