#!/bin/bash
# MN-major tf32 operand probe (SWIZZLE_128B_BASE32B)
timeout 120 tools/variants/umma_probe_mn 2>&1 | tee gpurun_out/umma_probe_mn.log
