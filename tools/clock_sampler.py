#!/usr/bin/env python
"""Light NVML sampler (separate process): one CSV line per period with the SM clock, its maximum and the
active clock-event reasons of one GPU.  Same information as `nvidia-smi --query-gpu=clocks.sm,... -lms`,
but only the three NVML calls that are needed (the full nvidia-smi query slowed a pipelined host loop by ~20 %).
usage: clock_sampler.py GPU_INDEX PERIOD_MS   (runs until terminated)"""
import sys
import time

import pynvml as nv

idx, period = int(sys.argv[1]), float(sys.argv[2]) / 1e3
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(idx)
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
R = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
out = sys.stdout
while True:
    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    out.write(f"{sm},{mx},{'|'.join(k for k, b in R.items() if bits & b)}\n")
    out.flush()
    time.sleep(period)
