cd /root/repo
B="python bench.py --no-cpu --no-probe --no-consumer --no-check --steps 4 --warmup 2"
for f in 45 65 75; do $B --opt flush_pct=$f > gpurun_out/r02f_flush$f.json 2>/dev/null; done
for v in 0 4; do $B --no-e2e --opt l2s_variant=$v > gpurun_out/r02f_l2v$v.json 2>/dev/null; done
$B --no-e2e --workload c3k63 --opt page_threads=640 > gpurun_out/r02f_k63_p640.json 2>/dev/null
$B --no-e2e --workload c3 --opt l2s_variant=1 > gpurun_out/r02f_c3_v1.json 2>/dev/null
$B --no-e2e --workload c3 --opt l2s_variant=4 > gpurun_out/r02f_c3_v4.json 2>/dev/null
echo done
