#!/bin/bash
# Round-2 opening GPU call: host diagnostics (why the CPU arm moved 5.9x between boxes), the K1/K3 knob sweep that round 1 left
# unmeasured, and ncu --set full captures of the SHIPPED K1 (128^3), K2 (256^3 field) and K3 (128^3) at the bench configurations.
#   gpurun --timeout 1500 -- 'bash tools/r2_call1.sh'
O=gpurun_out
mkdir -p $O
{
  echo "== host"; nproc; lscpu | grep -E 'Model name|Socket|Core|Thread|NUMA|^CPU\(s\)'
  echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; echo "cpuset: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null)"
  grep -E 'Cpus_allowed_list' /proc/self/status; free -g | head -2; uptime
  cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -6
} > $O/r2a_host.txt 2>&1

# --- CPU arm under different thread settings (reference arm = oracle/_ref on the host cores)
for cfg in "DG_CPU_THREADS=128 DG_OMP_PROC_BIND=false" "DG_NOP=1" "DG_CPU_THREADS=128" "DG_CPU_THREADS=32" "DG_NOP=2"; do
  echo "== $cfg" >> $O/r2a_cpuarm.txt
  env $cfg timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d['cpu_baseline']; print(round(d['value']/1e6,2),'Mnodes/s mean;', [round(t,2) for t in d['timing']['per_step_s']], 's/step;', c['cores'],'threads', c.get('omp'), 'quota', c.get('cgroup_cpu_quota'), 'phys', c.get('physical_cores'), c['mode'])" >> $O/r2a_cpuarm.txt
done
cat /sys/fs/cgroup/cpu.stat 2>/dev/null | head -6 >> $O/r2a_host.txt

# --- knob sweep (K1-only bench at 128^3; parity first)
for so in build/variants/*.so; do
  n=$(basename $so .so)
  extra="--no-density --no-target"; tests="tests/test_gpu_k1_sdf.py"
  case $n in
    k3div) extra="--no-target"; tests="tests/test_gpu_k3_density.py";;
    base)  extra=""; DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --mesh torus --no-interp --no-cpu --no-e2e --no-real --no-density --no-target 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('base TORUS K1 128^3', round(d['ms_per_step'],2),'ms (round 1: 54.9)')";;
    cost|fastdiv) extra="--no-density";;
  esac
  ok=$(DISCREGRID_B200_LIB=$PWD/$so timeout 300 python -m pytest $tests -m gpu -q -x -k "not full_size" 2>&1 | tail -1)
  echo "$n parity: $ok"
  DISCREGRID_B200_LIB=$PWD/$so timeout 300 python bench.py --steps 5 --warmup 3 --no-interp --no-cpu --no-e2e --no-real $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('target_config') or {}; k=(d.get('density_map') or {})
print('$n', 'K1 128^3', round(d['ms_per_step'],2),'ms', round(d['value']/1e6,1),'Mnodes/s | target', round(t.get('ms_per_step',0),1),'ms | K3', round(k.get('ms',0),1),'ms')"
done > $O/r2a_sweep.txt 2>&1

# --- ncu --set full of the shipped kernels at the bench configurations (numbers under ncu are never bench values)
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:sdf_sample_nodes -s 1 -c 1 -f -o $O/r2a_k1_128 python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target --no-density > $O/r2a_ncu_k1.log 2>&1
timeout 400 $NCU -k regex:density_map -c 1 -f -o $O/r2a_k3_128 python bench.py --steps 1 --warmup 1 --no-interp --no-cpu --no-e2e --no-real --no-target > $O/r2a_ncu_k3.log 2>&1
timeout 400 $NCU -k regex:interpolate_kernel -s 2 -c 2 -f -o $O/r2a_k2_256 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-real --no-target --no-density > $O/r2a_ncu_k2.log 2>&1
ls -la $O
