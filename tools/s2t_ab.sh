T="timeout -s KILL 200"
$T python tools/s2_rate.py 600:592 1000:296 1500:148 2000:148 2>&1 | grep -v Warning
$T python tools/s2_rate.py 3375:148 2>&1 | grep -v Warning
SCQED_DTC_GP=1 $T python tools/s2_rate.py 1000:296 1500:148 2>&1 | grep -v Warning
