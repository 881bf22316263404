# Does a running `nvidia-smi --query-gpu=... -lms P` slow a kernel that is being timed?  Config 2's scan, 200 and 20 steps.
Q="index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
Q2="index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
run() { python tools/step_count_probe.py | grep -E "idle   0 ms  n=  (20|200) " | tr -s ' ' | sed "s/^/$1: /"; }
run "no sampler"
for P in 100 200 500 1000; do
  nvidia-smi --query-gpu=$Q --format=csv,noheader,nounits -lms $P -i 0 > /dev/null & S=$!
  sleep 1.5; run "sampler every $P ms"; kill $S; wait $S 2>/dev/null
done
nvidia-smi --query-gpu=$Q2 --format=csv,noheader,nounits -lms 100 -i 0 > /dev/null & S=$!
sleep 1.5; run "short query every 100 ms"; kill $S; wait $S 2>/dev/null
