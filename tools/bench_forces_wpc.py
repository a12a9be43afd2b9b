import sys
sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo')
from bench_forces import run
for w in (1,2):
    run("USA_Lanker-2_18_T-1_LF", 50, 8192, warps_per_cta=w)
    run("ZAM_Over-1_1_LF", 30, 1024, warps_per_cta=w)
    run("ZAM_Over-1_1_LF", 30, 32768, warps_per_cta=w)
    run("ZAM_Over-1_1_CA", 30, 4096, warps_per_cta=w)
