cd /root/repo
one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['rescanned_streams'])"; }
for cfg in "RFB_RING_CAP=32" "RFB_RING_CAP=16 RFB_HOT_ROWS=333" "RFB_RING_CAP=32 RFB_HOT_ROWS=110"; do for mix in adv wsplice uniform; do echo -n "$cfg $mix "; env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one; done; done
for cfg in "RFB_RING_CAP=32 RFB_HOT_ROWS=110"; do for mix in wmix whi wlo; do echo -n "$cfg $mix "; env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one; done; done
for cfg in "RFB_RING_CAP=32" "RFB_RING_CAP=16 RFB_HOT_ROWS=333"; do echo "$cfg l7"; env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --ruleset l7_filter 2>/dev/null | one; echo "$cfg config5"; env $cfg python tools/dev/config5_bench.py 262144 2>&1 | tail -2; done
