timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1"
P='import json,sys; d=json.loads(sys.stdin.read()); print("%s value %.1f M/s  ms %.2f frac %.3f acc %.4f" % (sys.argv[1], d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["accept_rate"]))'
$B 2>/dev/null | python -c "$P" fold46
SRHMC_CHAIN_WCUT_BITS=50 $B 2>/dev/null | python -c "$P" fold50
SRHMC_LIB=$PWD/hmc_stellar_toy_model_b200/libstellar_rhmc_nofold.so $B 2>/dev/null | python -c "$P" nofold46
$B --workload c5 2>/dev/null | python -c "$P" c5_t3
SRHMC_LIB=$PWD/hmc_stellar_toy_model_b200/libstellar_rhmc_t4.so $B --workload c5 2>/dev/null | python -c "$P" c5_t4
