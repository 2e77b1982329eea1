for d in 0 1000 2500 5000 10000; do echo "== stagger $d"; QCS_CUDA_STAGGER_NS=$d timeout 200 python scripts/quick_bench.py 28 ldg8 2>&1 | grep -E "rz_all|h_all|qft:|random"; done
