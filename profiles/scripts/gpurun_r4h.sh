# round 2, call 4h (1 GPU): few-coil owner-tile spread, warps per CTA A/B (B2N_OPT_ADJ_OWNED 7 = one warp, 1 = four warps / 40 regs, 8 = four warps / 32 regs)
mkdir -p gpurun_out
timeout 600 python profiles/scripts/adj_variants.py cfg1 cfg2 --variants= --caps=128 --owned=7,1,8 --coils=1,2,4 > gpurun_out/r4h_few_coils.log 2>&1
cat gpurun_out/r4h_few_coils.log | tail -30
