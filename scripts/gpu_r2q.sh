#!/bin/bash
# EKF rewrite: parity tests + phase profile + config 3 record
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "ekf or dob or ampc or tick or rls" > $O/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2q_pytest.log
BR2_VARIANT=prof timeout 300 python scripts/ekf_phase_profile.py > $O/r2q_ekf_phase.json 2> $O/r2q_ekf.err; echo "ekf rc=$?"; cat $O/r2q_ekf_phase.json; tail -3 $O/r2q_ekf.err
timeout 300 python scripts/ekf_phase_profile.py 2>&1 | head -3
