cd /root/repo
one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['rescanned_streams'])"; }
for cfg in "RFB_RING_CAP=32" "RFB_RING_CAP=16 RFB_HOT_ROWS=333" "RFB_RING_CAP=16 RFB_HOT_ROWS=120" "RFB_RING_CAP=16"; do for mix in wmix whi wlo; do echo -n "$cfg $mix "; env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one; done; done
