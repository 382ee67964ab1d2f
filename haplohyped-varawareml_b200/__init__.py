"""haplo-b200: the B200-native VCF -> tensor hot path of HaploHyped-VarAwareML.

Layout (only what the hot path needs):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/haplo_b200.h)
  capi.py          ctypes binding of libhaplo_b200.so (fails loudly if the library is missing)
  parse_vcf*.so    the reference-facing pybind11 module (built in-tree by build.py)
  vcf_to_h5.py     host-side mirror of src/haplohyped/vcf_to_h5.py (converter + click CLI)
  h5_reader.py, common_utils.py, haplotype_dataset.py   mirrors of src/utils, src/datasets
"""
__version__ = "0.1.0"
