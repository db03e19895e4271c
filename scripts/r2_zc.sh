# zero-copy transfer sweep (development): mode:up:down
python scripts/e2e_host_overhead.py 2>&1 | tail -1 | cut -c1-110
for c in ${COMBOS:-1:32:32 1:48:32 1:64:32 1:32:16 1:32:24 1:32:48 1:24:24 1:48:48 2:32:32 2:32:24 2:32:48 3:32:32 3:48:32 3:64:32}; do
  m=$(echo $c | cut -d: -f1); u=$(echo $c | cut -d: -f2); d=$(echo $c | cut -d: -f3)
  TROLLOUT_ZEROCOPY=$m TROLLOUT_ZC_UP=$u TROLLOUT_ZC_DOWN=$d python scripts/e2e_host_overhead.py 2>&1 | tail -1 | cut -c1-110 | sed "s/^/mode $m up $u down $d /"
done
