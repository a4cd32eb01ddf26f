#!/bin/bash
# last pass of round 2: entry-point tests with the final block sizing, I/O legs on the 2 M-tet and 10 M-tet meshes
set -u
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_cli.py tests/test_gpu_wss_matrix.py -m gpu -q > $OUT/pytest_cli.log 2>&1; echo "pytest cli rc=$? $(tail -1 $OUT/pytest_cli.log)"
timeout 200 python bench.py --workload aneurysm_p1 --snapshots 127 --steps 5 --no-cpu-baseline --no-other-workloads --io-gib 1.1 > $OUT/io_aneurysm_p1.json 2> $OUT/io_aneurysm_p1.err; echo "io aneurysm rc=$?"
timeout 300 python bench.py --workload vessel10m_p2 --snapshots 16 --steps 3 --no-cpu-baseline --no-other-workloads --io-gib 4 > $OUT/io_vessel10m_p2.json 2> $OUT/io_vessel10m_p2.err; echo "io 10m rc=$?"
python - $OUT <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], f'value {d["value"]/1e9:.2f} G/s e2e {d["e2e"]["value"]/1e9:.3f} G/s')
        if "io" in d: print("   io:", {k:(round(v["open_to_result_s"],3), round(v["file_read_s"],3), round(v["gbs"],2)) for k,v in d["io"].get("hdf5_to_device",{}).items()}, d["io"].get("entry_point",{}).get("breakdown"), d["io"].get("entry_point",{}).get("total_s"), d["io"].get("snapshots"), d["io"].get("bytes"), d["io"].get("error"))
    except Exception as e:
        print(f, "no line", e)
PY
