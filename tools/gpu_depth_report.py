"""gpurun tool: per-layer parity table of a full-depth step (tests/depth_parity.py) -> gpurun_out/depth_parity_<tag>.{json,md}.
usage: python tools/gpu_depth_report.py [c2|c3|c4|c5] [forced] [L=<layers>]"""
import dataclasses
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import __graft_entry__ as g

    g.build()
    from bya_b200.synth import CONFIGS
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from depth_parity import depth_parity, format_report

    name = next((a for a in sys.argv[1:] if a in CONFIGS), "c2")
    forced = "forced" in sys.argv[1:]
    cfg = CONFIGS[name]
    for a in sys.argv[1:]:
        if a.startswith("L="):
            cfg = dataclasses.replace(cfg, num_layers=int(a[2:]))
    res = depth_parity(cfg, forced_masks=forced)
    tag = f"{name}_{'forced' if forced else 'soft'}_L{cfg.num_layers}"
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(res, open(os.path.join(out, f"depth_parity_{tag}.json"), "w"), indent=1)
    md = format_report(res)
    open(os.path.join(out, f"depth_parity_{tag}.md"), "w").write(md + "\n\n" + json.dumps(res["seconds"]) + "\n")
    print(md[-1500:])
    print(res["output"], res["seconds"])


if __name__ == "__main__":
    main()
