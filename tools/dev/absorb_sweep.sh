cd /root/repo
one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['matches_per_step_rank0'], d['image']['n_sticky'])"; }
for a in 12 6 3 0; do for mix in wmix whi wlo; do echo -n "absorb=$a $mix "; RFB_DFA_ABSORB=$a timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --mix $mix 2>/dev/null | one; done; done
