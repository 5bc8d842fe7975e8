"""Take the C5 checksums out of a bench line (the default run's extra.c5, or a --workload c5 line) measured on ONE GPU and
write tests/golden/c5_field_checksum.json (what every later multi-GPU run and tests/test_gpu_properties.py compare with)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
assert d["n_gpus"] == 1
pc = (d["extra"].get("c5") or d["extra"])["parity_checked"]
assert int(pc["settled"]) == 241537616, pc
out = {"checksum_g": pc["checksum_g"], "checksum_pred": pc["checksum_pred"], "settled": 241537616,
       "what": "position-weighted sums mod 2^64 (nsfs.sharded_field.field_checksum) of the fp32 g plane (the exact Q17.15 fixed-point "
               "solution, converted once at the end) and the pred plane of the C5 field (16384x16384, p_occ 0.10, seed 11, source = centre), "
               "single-GPU solve; the solve is exact integer arithmetic with a unique fixed point, so every scheduler and every sharding "
               "must reproduce them"}
for path in (os.path.join(ROOT, "tests", "golden", "c5_field_checksum.json"), os.path.join(ROOT, "gpurun_out", "c5_field_checksum.json")):
    json.dump(out, open(path, "w"), indent=1)
print(out)
