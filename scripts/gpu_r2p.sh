#!/bin/bash
# host-path overhead breakdown + EKF phase profile
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/host_overhead_probe.py > $O/r2p_host_overhead.json 2> $O/r2p_host.err; echo "host rc=$?"; cat $O/r2p_host_overhead.json; tail -3 $O/r2p_host.err
BR2_VARIANT=prof timeout 300 python scripts/ekf_phase_profile.py > $O/r2p_ekf_phase.json 2> $O/r2p_ekf.err; echo "ekf rc=$?"; cat $O/r2p_ekf_phase.json; tail -3 $O/r2p_ekf.err
