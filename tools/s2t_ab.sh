# eigh_batched matrices/s by size through tools/s2_rate.py.  Variants by environment: SCQED_DENSE_TILED = 1 (register kernel) /
# 2 (streamed kernel), SCQED_DTC_NB = 8 / 4 / 2, SCQED_DTC_YS = 4 / 1, SCQED_DTC_NST (ring, units of 16 KB), SCQED_DTC_GP = 0 / 1
# (panels in shared / global memory), SCQED_DTC_TWO = 1 (two CTAs per SM).  Sizes >= 2900 in a process of their own (see DESIGN).
T="timeout -s KILL 200"
$T python tools/s2_rate.py 256:2072 300:2072 400:1184 500:1184 600:592 800:296 1000:296 1500:148 2000:148 2500:148 2>&1 | grep -v Warning
$T python tools/s2_rate.py 3000:148 2>&1 | grep -v Warning
$T python tools/s2_rate.py 3375:148 2>&1 | grep -v Warning
