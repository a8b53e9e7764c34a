"""Copies the reference's own site definitions (input CSVs) and golden OUTPUT tables into
tests/golden/ so that the tests can run on the GPU box, where /root/reference does not exist.

Run once in the build container:  python tests/golden/make_fixtures.py
Only data files are copied (inputs + golden result tables); no reference source code.
The inline golden tables of test/01-unit-tests.jl (:907-945 water partitioning, :951-1014 delta
profiles) are transcribed into golden_inline.json by parsing that test file.
"""
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

INPUT_SETS = {
    "DAV2020-full": "examples/DAV2020-full",
    "DAV2020-bare-minimum": "examples/DAV2020-bare-minimum",
    "BEA2016-reset-FALSE-input": "examples/BEA2016-reset-FALSE-input",
    "isoBEA2016-reset-FALSE-input": "examples/isoBEA2016-reset-FALSE-input",
    "infiltrationSaturationINFEXP1": "examples/infiltrationSaturationINFEXP1",
}
# only the files of one prefix are taken from this directory (root-model unit tests, test/01-unit-tests.jl:384-421)
HAMMEL = ("test/test-assets/Hammel-2001/input-files-ISO", "Hammel_loam-NLayer-27-RESET=FALSE")
GOLDEN_OUTPUTS = [
    "examples/DAV2020-full_u_sol_referenceu_mm.csv",
    "examples/DAV2020-full_u_sol_referenceu_δ.csv",
    "examples/DAV2020-full_u_sol_referenceu_belowground.csv",
    "examples/INFEXP1_fluxes_referencedf.csv",
    "examples/INFEXP1-modified_fluxes_referencedf.csv",
    "examples/DAV2020-full-modified_fluxes_referencedf.csv",
]


def parse_matrix(block: str):
    rows = []
    for line in block.strip().splitlines():
        line = line.split("#")[0].strip().rstrip(";").strip()
        if not line:
            continue
        if ";" in line:
            parts = [p for p in line.split(";") if p.strip()]
        else:
            parts = [line]
        for p in parts:
            vals = [float("nan") if v == "NaN" else float(v) for v in p.split()]
            rows.append(vals)
    return rows


def main():
    for name, rel in INPUT_SETS.items():
        dst = os.path.join(HERE, "inputs", name)
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(REF, rel))):
            if f.endswith(".csv") and "oldFormat" not in f:
                shutil.copy(os.path.join(REF, rel, f), os.path.join(dst, f))
    dst = os.path.join(HERE, "inputs", "Hammel-2001")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, HAMMEL[0]))):
        if f.startswith(HAMMEL[1]) and f.endswith(".csv"):
            shutil.copy(os.path.join(REF, HAMMEL[0], f), os.path.join(dst, f))
    # examples/isoBEA2016-reset-FALSE-input/*_param.csv is in a pre-0.9 format: it lacks 7 keys that
    # read_path_param (src/func_read_inputData.jl:978-1059) requires, so the reference itself cannot load
    # it.  The fixture takes the missing rows from the BEA2016-reset-FALSE param file (same site, same year).
    iso_p = os.path.join(HERE, "inputs", "isoBEA2016-reset-FALSE-input", "isoBEA2016-reset-FALSE_param.csv")
    bea_p = os.path.join(HERE, "inputs", "BEA2016-reset-FALSE-input", "BEA2016-reset-FALSE_param.csv")
    have = {l.split(",")[0].strip() for l in open(iso_p)}
    extra = [l for l in open(bea_p) if l.split(",")[0].strip() not in have and not l.startswith("###")]
    lines = open(iso_p).read().rstrip("\n").split("\n")
    with open(iso_p, "w") as fh:
        fh.write("\n".join(lines) + "\n" + "### rows added from BEA2016-reset-FALSE_param.csv (see make_fixtures.py),NA\n" + "".join(extra))
    os.makedirs(os.path.join(HERE, "outputs"), exist_ok=True)
    for rel in GOLDEN_OUTPUTS:
        shutil.copy(os.path.join(REF, rel), os.path.join(HERE, "outputs", os.path.basename(rel)))
    txt = open(os.path.join(REF, "test/01-unit-tests.jl"), encoding="utf-8").read()
    inline = {}
    for key in ("reference_daily_check", "reference_monthly_check", "reference_yearly_check",
                "reference_isotopePlotPermutedDims"):
        m = re.search(key + r"\s*=\s*\[(.*?)\]", txt, flags=re.S)
        inline[key] = parse_matrix(m.group(1))
    inline["_source"] = "test/01-unit-tests.jl:907-945, :951-1014 (LWFBrook90.jl v0.9.9)"
    json.dump(inline, open(os.path.join(HERE, "golden_inline.json"), "w"), indent=0)
    print({k: (len(v), len(v[0])) for k, v in inline.items() if not k.startswith("_")})


if __name__ == "__main__":
    main()
