python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_t8_tests.log
python tools/gpu_ab.py build_variants/libsf_base.so build_variants/libsf_cur.so build_variants/libsf_rmw.so build_variants/libsf_oldatan.so > gpurun_out/r02_t8_ab.log 2>&1
for v in bt_cur bt_rmw bt_oldatan; do SF_B200_LIB=build_variants/libsf_$v.so python tools/gpu_barrier_timing.py; done > gpurun_out/r02_t8_barrier.log 2>&1
