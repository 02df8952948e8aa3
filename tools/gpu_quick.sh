cd /root/repo
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; }
run late A=1
run inline RF_GLOBAL_WGRAD_INLINE=1
run late A=1
run inline RF_GLOBAL_WGRAD_INLINE=1
