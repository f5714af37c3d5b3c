T="timeout -s KILL 200"
$T python tools/s2_rate.py 2900:148 3000:148 3100:148 3375:148 3375:148 2>&1 | grep -v Warning
