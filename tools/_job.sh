for cap in 0 4; do
echo "HOH_DEC_CAP=$cap"
HOH_DEC_CAP=$cap python tools/jitter_probe.py 8192 256 256 4 12
HOH_DEC_CAP=$cap python tools/jitter_probe.py 64 3840 2160 2 8
done
