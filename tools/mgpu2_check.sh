# Two-GPU check (gpurun --gpus 2 -- bash tools/mgpu2_check.sh): the z-slab parity worker, then the bench line at N = 2.
python -c "import torch"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_worker.py 2>&1 | grep -i "pipelined\|tall\|ALL OK\|FAILED\|MISMATCH\|Error" | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --newton-steps 3 --rb-strong 1 2>&1 | grep '^{' > gpurun_out/bench_final_2gpu.json; echo "rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/bench_final_2gpu.json'):
    d = json.loads(l); n = d['newton']; r = d.get('rb_strong') or {}
    print('asm %.4f ms  e2e %.3f ms  spmv %.4f  newton %.1f ms %s  rb_strong asm %s newton %s  parity_slabs %s' % (
        d['ms_per_step'], d['e2e']['ms_per_step'], d['spmv']['ms'], n['ms_per_step'], n['krylov_iterations'],
        r.get('assembly_ms'), r.get('newton_ms_per_step'), d.get('parity_slabs')))
P
