#!/bin/bash
# Probe of the GPU box: is SimpleITK (the reference's N4 dependency, requirements.txt:12) present or
# installable there, and what does the host side look like (NUMA nodes, cores, PCIe topology)?
out=gpurun_out/probe_box.txt
{
echo "== date"; date -u
echo "== python -c 'import SimpleITK'"; python -c "import SimpleITK as s; print(s.Version())" 2>&1 | tail -2
echo "== python -c 'import itk'"; python -c "import itk; print(itk.Version.GetITKVersion())" 2>&1 | tail -1
echo "== pip download SimpleITK==2.3.1 (index)"; timeout 40 python -m pip download SimpleITK==2.3.1 -d /tmp/sitk --no-deps 2>&1 | tail -3
echo "== pip download --no-index --find-links /opt/wheelhouse"; python -m pip download SimpleITK==2.3.1 --no-index --find-links /opt/wheelhouse -d /tmp/sitk 2>&1 | tail -2
echo "== find / -iname '*simpleitk*' '*N4Bias*'"; find / \( -iname '*simpleitk*' -o -iname '*N4Bias*' \) -not -path '/proc/*' -not -path '*/repo/*' -not -path '/tmp/graft*' 2>/dev/null | head
echo "== nproc / lscpu"; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" 
echo "== numactl"; numactl -H 2>&1 | head -20
echo "== /sys numa nodes"; ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
echo "== nvidia-smi topo"; nvidia-smi topo -m 2>&1 | head -30
echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv
for g in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo "numa_node of $g: $(cat /sys/bus/pci/devices/$g/numa_node 2>/dev/null)"; done
echo "== free -g"; free -g | head -2
} > $out 2>&1
cat $out
