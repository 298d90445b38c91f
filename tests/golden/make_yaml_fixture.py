"""Golden fixture of the reference's shipped config surface: parses /root/reference/Config/*.yaml (the four files the entry
scripts are launched with, Script/train_stage{1,2,3}.sh, Script/eval scripts) and stores them verbatim as JSON, so that the CPU
tests can check that every shipped yaml is accepted UNCHANGED by lhrs_bot_b200 without /root/reference being present.

    python tests/golden/make_yaml_fixture.py      # writes tests/golden/shipped_yamls.json
"""
import glob
import json
import os

import yaml

REF = "/root/reference/Config"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shipped_yamls.json")

if __name__ == "__main__":
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, "*.yaml"))):
        with open(path) as f:
            out[os.path.basename(path)] = yaml.safe_load(f)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, {k: (v.get("stage"), v.get("dtype"), v.get("bits"), v["lora"]["enable"], v.get("tune_rgb_pooler")) for k, v in out.items()})
