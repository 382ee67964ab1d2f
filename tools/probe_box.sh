#!/bin/bash
# Probe the GPU box for the third-party libraries the reference's path uses.
# Output: gpurun_out/probe_box.txt
out=gpurun_out/probe_box.txt
mkdir -p gpurun_out
{
echo "== date: $(date -u)"; echo "== host: $(uname -a)"
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv,noheader
echo "== python modules"
for m in h5py hdf5plugin blosc2 blosc b2h5py tables numcodecs zarr pysam cyvcf2 polars lz4; do
  python -c "import $m; print('$m: present', getattr($m,'__version__',''))" 2>/dev/null || echo "$m: absent"
done
echo "== shared libraries (ldconfig)"
ldconfig -p | grep -i -E "hts|hdf5|blosc|lz4|zstd|libz\." || true
echo "== files"
find / -xdev \( -name "libhts*" -o -name "libblosc*" -o -name "libhdf5*" -o -name "htslib" -o -name "hts.h" -o -name "blosc*.h" -o -name "hdf5.h" \) -not -path "/proc/*" 2>/dev/null | head -20
echo "== tools"
for t in bgzip tabix bcftools h5dump h5ls; do command -v $t || echo "$t: absent"; done
echo "== /root/reference: $(ls -d /root/reference 2>&1)"; echo "== baseline/_ref: $(ls -d baseline/_ref 2>&1)"
echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "h5|hdf|blosc|pysam|cyvcf|polars|lz4|tables" || echo "(no matching wheels)"
echo "== cpus: $(nproc)  mem: $(free -g | awk '/Mem/{print $2}') GB"; numactl -H 2>/dev/null | head -5 || lscpu | grep -i numa
nvidia-smi topo -m 2>/dev/null | head -20
} > $out 2>&1
cat $out
