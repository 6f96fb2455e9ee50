for r in 1 2 4; do echo "rows=$r"; NBEST_LN_FWD_ROWS=$r timeout -s KILL 120 python profiles/ln_micro.py | grep -E "ln_fwd|copy"; done
