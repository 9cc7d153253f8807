for v in ${VARIANTS:-1 0}; do echo "=== DANET_TC_VARIANT=$v"; DANET_TC_VARIANT=$v timeout 300 python tools/tc_layers.py ${MODE:-all} 2>&1 | tail -${TAILN:-23} | cut -c1-100; done
