"""Mirror of the part of magpurify2/tools.py that sits on the hot paths: the native re-exports
(tools.py:39-42) and the scoring helpers (tools.py:283-351), same names and argument meaning,
computed by libmp2 on the GPU."""
import numpy as np

from . import _scoring
from ._composition import get_gc, get_tnf  # noqa: F401  (tools.py:40)
from ._coverage import get_bam_coverages  # noqa: F401  (tools.py:41)


def get_weighted_median(data, weights):
    """Weighted median of `data` (tools.py:283-311): sort by (value, weight); a single weight above
    half the total wins; an exact split returns the mean of the two straddling values."""
    data = np.array(data, dtype=np.float32).squeeze().reshape(-1)
    weights = np.array(weights).squeeze().reshape(-1)
    _s, med = _scoring.log_ratio_scores_batch(data, weights, [0, len(data)], 2.0, min_rows=1, want_medians=True)
    return med[0, 0]


def get_log_ratio_scores(data, weights, base, pseudo=1e-5, min=0.0, max=1.0):
    """1 - |log_base((x + pseudo) / (weighted median + pseudo))|, averaged over columns for 2-D
    input, clipped to [min, max] (tools.py:314-351)."""
    data = np.asarray(data, dtype=np.float32)
    return _scoring.log_ratio_scores_batch(data, weights, [0, data.shape[0]], base, pseudo=pseudo,
                                           clip=(min, max), min_rows=1)


# ---- helpers either side of the hot paths (same names / behaviour as magpurify2/tools.py) ----
import hashlib
import logging
import os
import struct
import sys
import zlib
from enum import Enum, auto

logger = logging.getLogger("timestamp")


class Compression(Enum):  # tools.py:45-49
    bzip2 = auto()
    gzip = auto()
    xz = auto()
    noncompressed = auto()


def validade_input(value, parameter_name, interval):
    """A numeric option forced into [min(interval), max(interval)]; the user is told when that changes it
    (behaviour of tools.py:54-82)."""
    lo, hi = min(interval), max(interval)
    clamped = lo if value < lo else hi if value > hi else value
    if clamped != value:
        logger.warning(f"The value of the `{parameter_name}` parameter must be between {lo} and {hi}. "
                       f"Setting it to {clamped}.")
    return clamped


def check_output_directory(output_directory):  # tools.py:85-97
    if not output_directory.is_dir():
        logger.warning("Output directory does not exist. Creating it now.")
        output_directory.mkdir()


def is_compressed(filepath):  # tools.py:123-147
    with open(filepath, "rb") as fin:
        signature = fin.read(8)
    if signature[:2] == b"\x1f\x8b":
        return Compression.gzip
    if signature[:3] == b"BZh":
        return Compression.bzip2
    if signature[:7] == b"\xfd7zXZ\x00\x00":
        return Compression.xz
    return Compression.noncompressed


def get_file_signature(filepath):
    """SHA-256 of the last 128,000 bytes (tools.py:150-166; OSError on smaller files, as there)."""
    with open(filepath, "rb") as fin:
        fin.seek(-128000, os.SEEK_END)
        return hashlib.sha256(fin.read(128000)).hexdigest()


def _bgzf_first_block(filepath):
    with open(filepath, "rb") as fin:
        head = fin.read(18)
        if len(head) < 18 or head[:4] != b"\x1f\x8b\x08\x04":
            return b""
        xlen = struct.unpack("<H", head[10:12])[0]
        fin.seek(12)
        extra = fin.read(xlen)
        bsize, o = None, 0
        while o + 4 <= len(extra):
            slen = struct.unpack("<H", extra[o + 2:o + 4])[0]
            if extra[o:o + 2] == b"BC" and slen == 2:
                bsize = struct.unpack("<H", extra[o + 4:o + 6])[0] + 1
            o += 4 + slen
        if bsize is None:
            return b""
        comp = fin.read(bsize - 12 - xlen - 8)
    try:
        return zlib.decompress(comp, -15)
    except zlib.error:
        return b""


def is_sorted_bam(filepath):
    """First line of the decompressed stream contains SO:coordinate (tools.py:169-186)."""
    data = _bgzf_first_block(filepath)
    return b"SO:coordinate" in data.split(b"\n", 1)[0].strip()


def check_bam_files(bam_files):  # tools.py:189-207
    if not all(filepath.exists() for filepath in bam_files):
        logger.error("At least one of the supplied BAM files does not exist.")
        sys.exit(1)
    if not all(is_sorted_bam(filepath) for filepath in bam_files):
        logger.error(
            "At least one of the supplied BAM files is not sorted by coordinate "
            "or does not contain a valid header. You can sort it using `samtools sort`."
        )
        sys.exit(1)


def read_fasta(filepath):
    """Yields (id with '~' -> '_', description, sequence) per record (tools.py:210-240).
    gzip / bzip2 / xz detected by magic bytes."""
    from ._fasta import Fasta

    with Fasta(filepath) as fa:
        for i in range(len(fa.ids)):
            yield fa.ids[i], fa.descriptions[i], fa.sequence(i)


def get_tsv_coverages(filepath, contig_set=None):
    """Coverage table in a text file: contig name, then one coverage per sample, tab-separated, no header
    (behaviour of tools.py:243-280).  Returns (names str[n], float64[n, n_samples]); with `contig_set` only the
    rows of those contigs.  Every line is a data line ('#' is not a comment)."""
    names, rows, width = [], [], None
    with open(filepath) as fin:
        for line in fin:
            cells = line.rstrip("\n").split("\t")
            if width is None:
                width = len(cells)
                if width < 2:
                    logger.error("The tabular coverage file must have at least two columns.")
                    sys.exit(1)
            if not line.strip():
                continue
            if contig_set and cells[0] not in contig_set:
                continue
            names.append(cells[0])
            # short or non-numeric cells read as NaN, as numpy.genfromtxt reports them
            vals = [np.nan] * (width - 1)
            for k, cell in enumerate(cells[1:width]):
                try:
                    vals[k] = float(cell)
                except ValueError:
                    pass
            rows.append(vals)
    matrix = np.array(rows, dtype=float).reshape(len(names), (width or 2) - 1)
    return np.array(names, dtype=str), matrix


def write_module_output(score_list, score_output_file):
    """One TSV for a module: the attribute names of the first score object as header, then the rows every
    score object yields (format of tools.py:607-626)."""
    with open(score_output_file, "w") as fout:
        fout.write("\t".join(score_list[0].attributes) + "\n")
        fout.writelines("\t".join(str(cell) for cell in row) + "\n" for scores in score_list for row in scores)


def get_checkm_contamination(filepath):
    """{genome: contamination} from a CheckM tabular file (`checkm qa --tab_table`), None when the file is
    missing or lacks the 'Bin Id' / 'Completeness' / 'Contamination' columns (tools.py:629-670)."""
    import csv
    from pathlib import Path

    filepath = Path(filepath)
    if not filepath.exists():
        logger.warning("The supplied CheckM tabular file was not found. The default probability threshold value "
                       "will be used instead.")
        return None
    with open(filepath) as fin:
        reader = csv.reader(fin, delimiter="\t")
        fields = next(reader, [])
        if not {"Bin Id", "Completeness", "Contamination"}.issubset(fields):
            logger.warning("The supplied CheckM tabular file is not correctly formatted. The default probability "
                           "threshold value will be used instead.")
            return None
        bin_col, cont_col = fields.index("Bin Id"), fields.index("Contamination")
        return {row[bin_col]: float(row[cont_col]) for row in reader if row}


def write_filtered_genome(mag, mags_contaminants_dict, filtered_output_directory, suffix):
    """FASTA of the contigs of `mag` that were not flagged (tools.py:673-699): '>description', sequence wrapped
    at 70 columns; genomes absent from the dictionary are skipped."""
    import textwrap
    from pathlib import Path

    if mag.genome not in mags_contaminants_dict:
        return None
    name = f"{mag.genome}.{suffix}.fna" if suffix else f"{mag.genome}.fna"
    output_fna = Path(filtered_output_directory).joinpath(name)
    flagged = mags_contaminants_dict[mag.genome]
    with open(output_fna, "w") as fout:
        for contig, description, sequence in mag:
            if flagged[contig]:
                continue
            fout.write(f">{description}\n")
            if any(c in sequence for c in "- \t"):  # textwrap may break at hyphens / blanks: keep its behaviour
                fout.write(f"{textwrap.fill(sequence, 70)}\n")
            else:
                fout.write("\n".join(sequence[o:o + 70] for o in range(0, len(sequence), 70)) + "\n")
    return output_fna
