# round 2, call 6b: few-coil owner-tile spread with 32 visits per staging round (B2N_OPT_ADJ_OWNED 7 = 16 as before); 3-D gather with 16 warps: parity
mkdir -p gpurun_out
timeout 600 python profiles/scripts/adj_variants.py cfg1 cfg2 --variants= --caps=128 --owned=7,1 --coils=1,2,4 > gpurun_out/r6b_few_coils.log 2>&1
cat gpurun_out/r6b_few_coils.log | tail -14
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cases_match_reference or ordered_tiled or tiled_kernels_match or tiled_3d or cfg1 or cfg4 or edge_shapes" > gpurun_out/r6b_pytest.log 2>&1
tail -2 gpurun_out/r6b_pytest.log
