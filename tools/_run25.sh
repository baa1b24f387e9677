set -x
timeout 600 python tools/chunk_probe.py 10 32768,65536,131072,262144 8,16 2>&1 | grep -v Warning
