cd /root/repo
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; }
run default A=1
run side-2 RF_SIDE_PRIO=-2
run side-3 RF_SIDE_PRIO=-3
run default A=1
run side-2 RF_SIDE_PRIO=-2
