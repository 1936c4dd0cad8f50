#!/bin/bash
timeout 120 tools/variants/umma_ss_bw_probe 2>&1 | tee gpurun_out/umma_ss_bw_probe.log
