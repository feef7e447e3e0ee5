"""Host-side BAM input of `cmd coverage`: the library's BGZF/BAM decoder (csrc/bam.cu) behind the
names the reference uses (bam::IndexedReader / header / fetch, lib-coverage/src/lib.rs:533-547),
plus the sample-name rule of lib-coverage/src/lib.rs:36-64.

    with BamReader(path) as bam:
        bam.decode(min_unclipped=0.6)
        regions = bam.regions(contig_regex=r"^(chr)?\\d\\d?$")
        table = pipeline.coverage_run(regions, options)

Decoding is host work (zlib on threads) and is timed separately from the device path
(`decode_seconds`, `stats()`), as BASELINE.json's north_star asks.
"""
import ctypes as C
import re

import numpy as np

from . import _native as N


class BamError(N.CnvbError):
    pass


class BamReader:
    def __init__(self, path, threads=0):
        self._h = C.c_void_p()
        self.path = str(path)
        rc = N.lib().cnvb_bam_open(str(path).encode(), int(threads), C.byref(self._h))
        if rc != N.OK:
            msg = N.lib().cnvb_bam_error(self._h).decode() if self._h else "could not open %s" % path
            self.close()
            raise BamError(rc, msg)
        L = N.lib()
        n = L.cnvb_bam_n_refs(self._h)
        self.references = [L.cnvb_bam_ref_name(self._h, i).decode() for i in range(n)]
        self.lengths = [int(L.cnvb_bam_ref_len(self._h, i)) for i in range(n)]
        ln = C.c_uint64(0)
        p = L.cnvb_bam_header_text(self._h, C.byref(ln))
        self.header_text = C.string_at(p, ln.value).decode(errors="replace") if p and ln.value else ""
        self.decoded = False

    # -- the reference's sample-name rule: SM of the @RG lines, joined with ',' (lib.rs:36-64)
    def sample_names(self):
        """Sample list of the header, as samples_from_header / parse_line_rg build it
        (lib-shared/src/bam_utils.rs:43-87): per @RG line the LAST `ID:` and `SM:` fields, each the text
        between the first and the second ':' of a tab-separated field; lines lacking either are skipped;
        first occurrence order, no duplicates."""
        out = []
        for line in self.header_text.splitlines():
            if not line.startswith("@RG"):
                continue
            rg_id = sm = None
            for field in line.split("\t"):
                token = field.split(":")
                if len(token) >= 2:
                    if token[0] == "ID":
                        rg_id = token[1]
                    elif token[0] == "SM":
                        sm = token[1]
            if rg_id is not None and sm is not None and sm not in out:
                out.append(sm)
        return out

    def decode(self, min_unclipped=0.6, pinned=True):
        """One pass over the file: every record into the SoA columns of its contig."""
        rc = N.lib().cnvb_bam_decode(self._h, C.c_float(float(np.float32(min_unclipped))), 1 if pinned else 0)
        if rc != N.OK:
            raise BamError(rc, N.lib().cnvb_bam_error(self._h).decode())
        self.decoded = True
        return self

    def stats(self):
        counts = (C.c_uint64 * 4)()
        secs = (C.c_double * 3)()
        N.lib().cnvb_bam_stats(self._h, counts, secs)
        return dict(records=int(counts[0]), unplaced=int(counts[1]), compressed_bytes=int(counts[2]),
                    uncompressed_bytes=int(counts[3]), decode_seconds=float(secs[0]),
                    inflate_seconds=float(secs[1]), parse_seconds=float(secs[2]))

    @property
    def decode_seconds(self):
        return self.stats()["decode_seconds"]

    def fetch(self, tid, start=None, end=None):
        """The records of contig `tid` (name or index) as a ReadBatch over the decoder's buffers
        (valid while this reader is open and not decoded again).  With `start` / `end` (0-based,
        half-open): the records htslib's `fetch(tid, start, end)` yields, i.e. those overlapping the
        interval (pos < end and bam_endpos > start), in file order -- a selection from the decoded
        contig (records are in coordinate order: the candidates end at the first pos >= end)."""
        from .coverage import ReadBatch
        if not self.decoded:
            raise BamError(N.ERR_STATE, "BamReader.fetch before decode()")
        if isinstance(tid, str):
            tid = self.references.index(tid)
        soa = N.ReadSoa()
        rc = N.lib().cnvb_bam_ref_reads(self._h, int(tid), C.byref(soa))
        if rc != N.OK:
            raise BamError(rc, "no such reference: %r" % (tid,))
        n = int(soa.n)

        def col(ptr, ctype, dtype):
            if n == 0:
                return np.empty(0, dtype)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,))
            return a.view(dtype) if a.dtype != dtype else a

        b = ReadBatch(pos=col(soa.pos, C.c_int32, np.int32), end=col(soa.end, C.c_int32, np.int32),
                      mid=col(soa.mid, C.c_int32, np.int32), frag_end=col(soa.frag_end, C.c_int32, np.int32),
                      mapq=col(soa.mapq, C.c_uint8, np.uint8), flags=col(soa.flags, C.c_uint16, np.uint16))
        b._owner = self  # keep the buffers alive
        if start is not None or end is not None:
            lo_pos = 0 if start is None else int(start)
            hi = n if end is None else int(np.searchsorted(b.pos, int(end), side="left"))
            keep = np.flatnonzero(b.end[:hi] > lo_pos)
            b = ReadBatch(**{k: np.ascontiguousarray(getattr(b, k)[keep])
                             for k in ("pos", "end", "mid", "frag_end", "mapq", "flags")})
        return b

    def has_index(self):
        import os
        return os.path.exists(self.path + ".bai") or (self.path.endswith(".bam") and os.path.exists(self.path[:-4] + ".bai"))

    def fetch_indexed(self, tid, start, end, min_unclipped=0.6, pinned=False):
        """bam::IndexedReader::fetch(tid, start, end) + the record loop (lib-coverage/src/lib.rs:533-547) through
        the .bai index: only the blocks of the interval are read and inflated (cnvb_bam_fetch).  Returns the
        ReadBatch of the records overlapping [start, end) -- views of the decoder's columns of `tid`, valid until
        the next fetch / decode."""
        if isinstance(tid, str):
            tid = self.references.index(tid)
        rc = N.lib().cnvb_bam_fetch(self._h, int(tid), int(start), int(end), C.c_float(float(np.float32(min_unclipped))),
                                    1 if pinned else 0)
        if rc != N.OK:
            raise BamError(rc, N.lib().cnvb_bam_error(self._h).decode())
        self.decoded = True
        return self.fetch(tid)

    def regions(self, contig_regex=r"^(chr)?\d\d?$", sequences=None, targets=None):
        """The region loop's inputs (GenomeRegions::from_name_and_length over the contigs matching
        --contig-regex, lib-coverage/src/lib.rs:693-720): one pipeline.Region per matching contig."""
        from .pipeline import Region
        rx = re.compile(contig_regex)
        out = []
        for tid, (name, length) in enumerate(zip(self.references, self.lengths)):
            if rx.search(name):
                out.append(Region(name, length, self.fetch(tid),
                                  None if sequences is None else sequences.get(name),
                                  None if targets is None else targets.get(name)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            N.lib().cnvb_bam_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
