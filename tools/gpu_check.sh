#!/bin/bash
# Short check after a kernel change: parity tests, the CIGAR stage on HiFi-like records, the bench line.
TAG="${1:-run}"
O=gpurun_out
mkdir -p $O
timeout 200 python -m pytest tests -q -m gpu -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $O/${TAG}_pytest.log
timeout 100 python tools/cigar_bench.py --records 3000000 --mean-ops 31 --sigma 0.3 > $O/${TAG}_cigar_hifi.json 2> $O/${TAG}_cigar_hifi.err; echo "cigar exit $?"
timeout 200 python bench.py --steps 100 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
timeout 100 python tools/scale_check.py --gbp 3.1 --contigs 24 --coverage 30 --second-aligner > $O/${TAG}_scale_3gbp_hifi_2files.json 2> $O/${TAG}_scale3.err; echo "scale exit $?"
