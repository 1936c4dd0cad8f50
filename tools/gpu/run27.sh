timeout 60 tools/variants/umma_probe_bf16mn 2>&1 | tee gpurun_out/umma_probe_bf16mn.log
