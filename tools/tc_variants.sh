timeout ${TMO:-120} python tools/tc_layers.py ${MODE:-all} 2>&1 | tail -${TAILN:-23} | cut -c1-100
