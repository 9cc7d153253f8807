for v in ${PDLS:-0 1}; do echo "=== DANET_TC_PDL=$v"; DANET_TC_PDL=$v timeout 300 python tools/tc_layers.py ${MODE:-quick} 2>&1 | tail -${TAILN:-10} | cut -c1-100; done
