# round 2: two GPUs -- host and merge tests, smoke(), the torchrun bench (both arms)
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_host_gpu.py tests/test_shard_gpu.py -x -q > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?"; tail -2 $O/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err; echo "ref rc=$?"
python - <<PY
import json
d = json.loads(open("$O/bench_2gpu.json").read().strip().splitlines()[-1])
for k, b in {"primary": d, **d.get("workloads", {})}.items():
    r = b["roofline"]; e = b["e2e"]
    print(k, "value %.3f G/s" % (b["value"] / 1e9), "ms/step %.3f" % b["ms_per_step"], "scan %.3f path %.3f" % (r["frac"], r["path"]["frac"]),
          "e2e %.3f G/s h2d %.1f GB/s ceiling %.1f frac %.2f" % (e["value"] / 1e9, e["h2d_gbs"], e["h2d_ceiling_gbs"], e["h2d_frac_of_ceiling"]), b["clocks"]["samples"], b["clocks"]["reasons"])
print(d["verify_merge"])
PY
