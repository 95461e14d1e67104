python -c "import __graft_entry__ as g; g.smoke()"
timeout 100 python -m pytest tests/test_weno_gpu.py tests/test_slab_c_gpu.py tests/test_pent_part_gpu.py -q -m gpu -x 2>&1 | tail -2
