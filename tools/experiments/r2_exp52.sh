# is a batch of streams bound by the host's issue rate?
HOSTTIME=32 timeout 300 python tools/seg_only.py 2>&1 | grep "host issue"
