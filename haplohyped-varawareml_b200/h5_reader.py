"""Host-side mirror of the reference's src/utils/h5_reader.py (VCFH5Reader, :23-43).

Same class, method and error contract: fetch_genotypes(donor_id, chromosome) returns the whole
`donor_{id}/chr_{n}` record array (35-byte compound dtype) and raises KeyError when the group is
absent.  Repair R4: the dataset is read under the name the writer uses, `snp_data`
(the reference reader asks for `genotype`, which the writer never creates).

Container access goes through `container.open_h5`: h5py + hdf5plugin when they are importable,
otherwise this repo's minimal HDF5 reader; chunks are bare Blosc chunks (filter 32001) either way.
"""
from __future__ import annotations

import numpy as np

RECORD_DTYPE = np.dtype([("chrom", "S5"), ("start", np.uint32), ("stop", np.uint32),
                         ("ref", "S10"), ("alt", "S10"), ("phase1", np.int8), ("phase2", np.int8)])


class VCFH5Reader:
    def __init__(self, h5_file):
        from .container import open_h5
        self.h5_file = h5_file
        self.hdf5_file = open_h5(h5_file, "r")

    def fetch_genotypes(self, donor_id, chromosome):
        group_path = f"donor_{donor_id}/chr_{chromosome}"
        if group_path in self.hdf5_file:
            return self.hdf5_file.read_dataset(group_path + "/snp_data")
        raise KeyError(f"No data found for {group_path}")

    def stored_chunks(self, donor_id, chromosome):
        """The stored (compressed) chunks of donor_{id}/chr_{n}/snp_data, undecoded: (n_records, chunk_records, dtype,
        [bytes]) or None when the dataset is not stored that way.  What the dataset's device-resident store keeps."""
        group_path = f"donor_{donor_id}/chr_{chromosome}"
        if group_path in self.hdf5_file:
            return self.hdf5_file.stored_chunks(group_path + "/snp_data")
        raise KeyError(f"No data found for {group_path}")

    def close(self):
        self.hdf5_file.close()
