# GPU session script (round 2, #27): where a frame of the frontend loop (C5) spends its time: host profile + stage trace
timeout 300 python profiles/prof_c5_host.py 2>&1 | tail -32
PCR_TRACE=1 timeout 300 python profiles/prof_c5_host.py 2>&1 | grep -i "pcr" | tail -12
