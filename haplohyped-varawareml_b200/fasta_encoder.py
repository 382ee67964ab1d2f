"""Host-side mirror of the reference's reference-genome encoder (src/haplohyped/fasta_encoder.py): FASTA -> the
HDF5 file RandomHaplotypeDataset reads its windows from (`hdf5_reference_file`).

Same class / method names and click options (`--fasta --outdir --cores`, :189-192), same output path
(`{outdir}/reference_genome.h5`, :200), same chromosome list (chr1..chr22, :101), same `tmp_chrom_files/`
directory created and removed (:196,:208).  What is stored differs on purpose:

* the reference stores a 5-column one-hot per base (and its merge step overwrites every chromosome with the last,
  `/sequence/sequence`, :173-179); the dataset kernel (hb_hap.cu) gathers BASES and builds the one-hot on the fly
  for both haplotypes, so this writer stores one byte per base, dataset `chr{N}` -- the key the dataset asks for
  (haplotype_dataset.py:70 with repair R4) -- 5x smaller and readable in place;
* `encode_sequence` keeps the reference's contract (:63-78: str or |S1 array, upper-cased, non-ACGT -> N, one-hot
  with `encode_spec` columns) but runs on the GPU (kernel 5 through `onehot_windows`) and returns columns in
  `encode_spec` ORDER -- the reference sorts them alphabetically (:60), which disagrees with its own dataset.
"""
from __future__ import annotations

import gzip
import logging
import os
import shutil
from typing import Dict, Iterator, Tuple

import numpy as np

from .common_utils import parse_encode_dict
from .container import open_h5

logger = logging.getLogger(__name__)


def read_fasta(path: str) -> Iterator[Tuple[str, np.ndarray]]:
    """(name, bases as uint8 ASCII) per record of a FASTA file (plain or gzip); the name is the header up to the first
    white space, as pysam's FastaFile.fetch(chrom) keys it (:88-89)."""
    opener = gzip.open if path.endswith((".gz", ".bgz")) else open
    name, parts = None, []
    with opener(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    yield name, np.frombuffer(b"".join(parts), np.uint8)
                name, parts = line[1:].split()[0].decode() if line[1:].split() else "", []
            elif name is not None:
                parts.append(line.strip())
    if name is not None:
        yield name, np.frombuffer(b"".join(parts), np.uint8)


class ReferenceGenome:
    def __init__(self, fasta_file=None, encode_spec=None, hdf5_file=None, output_dir=None, device="cuda:0"):
        self.encode_spec = parse_encode_dict(encode_spec)
        self.output_dir = output_dir
        self.fasta_file = fasta_file
        self.hdf5_file = hdf5_file
        self.device = device
        self.genome: Dict[str, np.ndarray] = {}

    @staticmethod
    def parse_encode_list(encode_spec):
        """fasta_encoder.py:31-44, as bytes in spec order."""
        return [k.encode() for k in parse_encode_dict(encode_spec)]

    def encode_sequence(self, seq_data, ignore_case=True):
        """One-hot (L, C) float32 of a str / |S1 array, computed on the GPU."""
        from .haplotype_dataset import onehot_windows
        if isinstance(seq_data, str):
            arr = np.frombuffer(seq_data.encode("latin-1"), np.uint8)
        elif isinstance(seq_data, np.ndarray):
            arr = np.ascontiguousarray(seq_data.astype("|S1") if seq_data.dtype != "|S1" else seq_data).view(np.uint8)
        else:
            raise TypeError("Please input as string or numpy array!")
        if arr.size == 0:
            return np.zeros((0, len(self.encode_spec)), np.float32)
        return onehot_windows([arr], arr.size, self.encode_spec, ignore_case, self.device)[0].cpu().numpy()

    def load_genome_parallel(self, chromosomes=None):
        """FASTA -> {chrom: bases}.  (Nothing here is worth a thread pool: one sequential read of the file.)"""
        want = set(chromosomes) if chromosomes is not None else {f"chr{i}" for i in range(1, 23)}
        logger.info("Starting encoding of genome")
        for name, seq in read_fasta(self.fasta_file):
            if name in want:
                self.genome[name] = seq
                logger.info(f"Loaded chromosome {name}: {seq.size} bases")
        return self.genome

    def get_sequence(self, chrom, start, end):
        return self.genome[chrom][start:end].view("|S1")


class HDF5Handler:
    @staticmethod
    def save_to_hdf5(genome: Dict[str, np.ndarray], hdf5_file: str, backend=None):
        logger.info(f"Saving entire reference genome to {hdf5_file}")
        with open_h5(hdf5_file, "w", backend=backend) as f:
            for chrom, seq in genome.items():
                f.write_array(chrom, np.ascontiguousarray(seq).view("S1"))

    @staticmethod
    def load_from_hdf5(hdf5_file: str) -> Dict[str, np.ndarray]:
        out = {}
        with open_h5(hdf5_file, "r") as f:
            for chrom in f.keys():
                out[chrom] = np.ascontiguousarray(f.read_dataset(chrom)).view(np.uint8).reshape(-1)
        return out


def main(argv=None):
    import click

    @click.command()
    @click.option("--fasta", required=True, type=click.Path(exists=True), help="Path to reference genome FASTA file")
    @click.option("--outdir", required=True, type=click.Path(), help="Path to results save folder")
    @click.option("--cores", default=os.cpu_count(), type=int, help="Number of CPU cores to use")
    def _main(fasta, outdir, cores):
        output_dir = os.path.join(outdir, "tmp_chrom_files")
        os.makedirs(output_dir, exist_ok=True)
        ref_hdf5_file = os.path.join(outdir, "reference_genome.h5")
        try:
            ref_genome = ReferenceGenome(fasta_file=fasta, output_dir=output_dir)
            HDF5Handler.save_to_hdf5(ref_genome.load_genome_parallel(), ref_hdf5_file)
        finally:
            shutil.rmtree(output_dir, ignore_errors=True)
        logger.info(f"Reference genome HDF5 file created at {ref_hdf5_file}")

    return _main(args=argv, standalone_mode=argv is None)


if __name__ == "__main__":
    main()
