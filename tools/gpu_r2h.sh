#!/bin/bash
# final 1-GPU pass of round 2: suite, smoke, default bench + reference arm, launch-shape A/B, N=1 strong baseline, I/O legs
set -u
OUT=gpurun_out/r2h; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
timeout 120 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $OUT/smoke.log)"
( time timeout 900 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -4 $OUT/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; echo "ref rc=$?"
B="--no-cpu-baseline --no-io-leg --no-other-workloads"
VASP_B200_K2_OCC=3 timeout 300 python bench.py --workload avf_p2 --snapshots 256 --steps 10 $B > $OUT/occ3_avf_p2.json 2> $OUT/occ3_avf_p2.err; echo "occ3 avf rc=$?"
VASP_B200_K2_OCC=4 timeout 300 python bench.py --workload stenosis_p1 --steps 100 $B > $OUT/occ4_stenosis_p1.json 2> $OUT/occ4_stenosis_p1.err; echo "occ4 p1 rc=$?"
VASP_B200_K2_OCC=3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q > $OUT/pytest_occ3.log 2>&1; echo "pytest occ3 rc=$? $(tail -1 $OUT/pytest_occ3.log)"
VASP_B200_K2_OCC=4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q > $OUT/pytest_occ4.log 2>&1; echo "pytest occ4 rc=$? $(tail -1 $OUT/pytest_occ4.log)"
timeout 300 python bench.py --workload aneurysm_p1 --snapshots 2000 --scaling strong --steps 5 $B > $OUT/strong_aneurysm_p1_n1.json 2> $OUT/strong_aneurysm_p1_n1.err; echo "aneurysm n1 rc=$?"
timeout 300 python bench.py --workload aneurysm_p1 --snapshots 127 --steps 5 --no-cpu-baseline --no-other-workloads --io-gib 1.1 > $OUT/io_aneurysm_p1.json 2> $OUT/io_aneurysm_p1.err; echo "io aneurysm rc=$?"
timeout 600 python bench.py --workload vessel10m_p2 --snapshots 64 --steps 3 --no-cpu-baseline --no-other-workloads --io-gib 4 --e2e-snapshots 8 > $OUT/io_vessel10m_p2.json 2> $OUT/io_vessel10m_p2.err; echo "io 10m rc=$?"
python - $OUT <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        if d.get("impl")=="reference": print(f.split("/")[-1], "reference value %.3g"%d["value"]); continue
        print(f.split("/")[-1], f'step {d["ms_per_step"]:.4f} ms value {d["value"]/1e9:.2f} G/s e2e {d["e2e"]["value"]/1e9:.3f} G/s', {k: round(v["ms_per_launch"],4) for k,v in d["roofline"]["kernels"].items()})
        if "io" in d: print("   io:", {k:(round(v["open_to_result_s"],3), round(v["gbs"],2)) for k,v in d["io"].get("hdf5_to_device",{}).items()}, d["io"].get("entry_point",{}).get("breakdown"), d["io"].get("entry_point",{}).get("total_s"), d["io"].get("error"))
        for e in d.get("other_workloads",[]): print("   other:", e.get("workload"), e.get("error") or (round(e["value"]/1e9,2), round(e["ms_per_step"],4), round(e["e2e"]["value"]/1e9,3)))
    except Exception as e:
        print(f, "no line", e)
PY
