#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:channels_fwd -c 1 -s 1 -f -o $O/k44_pyr python tools/time_pyramid_warp.py > /dev/null 2>&1
ncu -i $O/k44_pyr.ncu-rep --page details 2>/dev/null | grep -E "Duration|Memory Throughput|DRAM Throughput|L1/TEX Hit|L2 Hit|Executed Ipc|Issue Slots|Achieved Occ|Theoretical Occ|Registers|Waves|Mem Busy|Max Bandwidth|Avg. Active Threads" | head -20
ncu -i $O/k44_pyr.ncu-rep --page raw 2>/dev/null | grep -E "smsp__average_warps_issue_stalled_(long|short|wait|lg|mio|math|barrier|sleeping|not_sel|no_inst).*ratio|dram__bytes_(read|write).sum |smsp__inst_executed.sum |l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum |lts__t_sectors_op_read.sum " | head -20
