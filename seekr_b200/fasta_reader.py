"""FASTA ingest with the reference's ``Reader`` interface (seekr/fasta_reader.py:9-109).

Parsing and 2-bit packing happen in the C++ packer of libseekr_b200 (host side, multi-threaded,
pinned output); this module only exposes the results with the reference's names and semantics:
``Reader(infasta).get_seqs()`` / ``get_headers()`` / ``get_lines()`` / ``get_data()``,
``supply_basic_header()`` and ``save()``.

``PackedFasta`` is what ``BasicCounter`` keeps instead of a list of Python strings: the packed
arrays go to the GPU as they are, and ``.seqs`` decodes a record to ``str`` only when indexed.
"""

import ctypes
import mmap
import os

import numpy as np

from . import _lib


def alphabet_lut(alphabet="AGTC"):
    """256-entry table byte -> digit (0..3) or 255, for the upper-cased sequence.

    The reference upper-cases every sequence (fasta_reader.py:55,62) and looks k-mers up in a map
    built from ``alphabet`` as given (kmer_counts.py:121-122), so a lower-case alphabet letter can
    never match; only 4-letter alphabets of distinct single-byte characters are packed in 2 bits.
    """
    if len(alphabet) != 4 or len(set(alphabet)) != 4:
        raise NotImplementedError(
            "seekr_b200 packs sequences in 2 bits per base: the alphabet must have 4 distinct letters, got %r"
            % (alphabet,))
    lut = np.full(256, 255, dtype=np.uint8)
    for digit, ch in enumerate(alphabet):
        code = ord(ch)
        if code > 255:
            raise NotImplementedError("non-latin-1 alphabet letter %r" % ch)
        if ch == ch.upper():
            lut[code] = digit
    return lut


_LOWER_TO_INVALID = bytes(0 if ord("a") <= b <= ord("z") else b for b in range(256))


def default_pack_threads():
    """Host threads for the packer: all cores, divided among the ranks of a torchrun launch on this node."""
    try:
        forced = int(os.environ.get("SEEKR_B200_PACK_THREADS", "0"))
    except ValueError:
        forced = 0
    if forced > 0:
        return min(forced, 256)
    cores = os.cpu_count() or 1
    try:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    except ValueError:
        local_world = 1
    return max(1, min(64, cores // max(1, local_world)))


class PackedFasta:
    """Owner of one SkrPacked handle plus the source text it was parsed from."""

    # attributes that need the COMPLETE record table; for a handle whose scan still runs in the background (large
    # texts, background=True) they are resolved on first use, which waits for the scan and raises what it found
    _TABLE_ATTRS = frozenset(("m", "nblocks", "total_bases", "slab_ptr", "slab_bytes", "off_codes", "off_mask",
                              "off_blk", "off_len", "max_length"))

    def __init__(self, handle, text=None, scanning=False):
        self._lib = _lib.load()
        self._h = handle
        self._text = text  # bytes-like (mmap or bytes) or None for skr_pack_sequences input
        if not scanning:
            self._resolve()

    def _resolve(self):
        lib, handle = self._lib, self._h
        m = int(lib.skr_packed_num_records(handle))
        if m < 0:  # the background scan found an error: raise it as the synchronous call would have
            _lib.check(lib.skr_packed_wait(handle))
            raise _lib.SeekrB200Error("background scan failed")
        self.m = m
        self.nblocks = int(lib.skr_packed_num_blocks(handle))
        self.total_bases = int(lib.skr_packed_total_bases(handle))
        self.slab_ptr = lib.skr_packed_slab(handle)
        self.slab_bytes = int(lib.skr_packed_slab_bytes(handle))
        base = self.slab_ptr or 0
        self.off_codes = (lib.skr_packed_codes(handle) or 0) - base
        self.off_mask = (lib.skr_packed_mask(handle) or 0) - base
        self.off_blk = (lib.skr_packed_block_offsets(handle) or 0) - base
        self.off_len = (lib.skr_packed_lengths(handle) or 0) - base
        # the longest record (the record table is final even while a background packer still runs)
        self.max_length = int(self.lengths.max()) if self.m else 0

    def __getattr__(self, name):
        # reached only when the attribute is not set yet: a handle that is still being scanned
        if name in PackedFasta._TABLE_ATTRS and self.__dict__.get("_h") is not None:
            self._resolve()
            return self.__dict__[name]
        raise AttributeError(name)

    @property
    def scanning(self):
        """True while the record count is not known yet (nothing has asked for it)."""
        return "m" not in self.__dict__

    def capacity(self):
        """Records and slab bytes the handle is laid out for (>= the final figures while the scan runs)."""
        lib = self._lib
        return int(lib.skr_packed_capacity_records(self._h)), int(lib.skr_packed_slab_bytes(self._h))

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def from_file(cls, path, alphabet="AGTC", pinned=False, nthreads=0, background=False):
        lib = _lib.load()
        lut = alphabet_lut(alphabet)
        size = os.path.getsize(path)
        with open(path, "rb") as handle:
            text = mmap.mmap(handle.fileno(), 0, access=mmap.ACCESS_READ) if size else b""
        return cls.from_buffer(text, alphabet, pinned, nthreads, _lut=lut, _lib_=lib, background=background)

    @classmethod
    def from_buffer(cls, text, alphabet="AGTC", pinned=False, nthreads=0, _lut=None, _lib_=None, background=False):
        """background=True returns once the text has been scanned (records, lengths, every error of the synchronous
        call); the code / mask words are packed by host threads behind the caller's back, in record order, and
        ``wait()`` (implied by the ``codes`` / ``mask`` views) or the streamed count path picks them up."""
        lib = _lib_ or _lib.load()
        lut = _lut if _lut is not None else alphabet_lut(alphabet)
        if nthreads <= 0 and len(text) > (1 << 22):
            nthreads = default_pack_threads()
            if background and nthreads >= 8 and "SEEKR_B200_PACK_THREADS" not in os.environ:
                # the packer runs beside the thread that drives the copy / count pipeline and the CUDA driver's own
                # threads: with 16 cores, two left to them finish earlier than all cores packing (17.4 against 18.8 ms,
                # profiles/r02_e2e_waves.txt); with 8 cores per rank one is the better trade (22.4 ms with 7 or 8
                # threads against 25.4 with 6: the pipeline is pack-bound there)
                nthreads -= 2 if nthreads >= 12 else 1
        n = len(text)
        if n:
            view = np.frombuffer(text, dtype=np.uint8)
            addr = ctypes.c_void_p(view.ctypes.data)
        else:
            addr = ctypes.c_void_p(0)
        out = ctypes.c_void_p()
        pack = lib.skr_pack_fasta_buffer_async if background else lib.skr_pack_fasta_buffer
        rc = pack(addr, n, ctypes.c_void_p(lut.ctypes.data), nthreads, int(pinned), ctypes.byref(out))
        _lib.check(rc)
        # large texts come back while the scan is still running: the record table is resolved on first use
        scanning = False
        if background:
            avail, fin = ctypes.c_int64(), ctypes.c_int()
            rc = lib.skr_packed_wait_scanned(out, 0, ctypes.byref(avail), ctypes.byref(fin))
            scanning = rc != _lib.SKR_OK or not fin.value
        obj = cls(out, text, scanning=scanning)
        obj._pending = bool(background)
        return obj

    @classmethod
    def from_sequences(cls, seqs, alphabet="AGTC", pinned=False, nthreads=0):
        """list[str] (e.g. ``BasicCounter.seqs`` assigned by hand)."""
        lib = _lib.load()
        lut = alphabet_lut(alphabet)
        offs = np.zeros(len(seqs) + 1, dtype=np.int64)
        if len(seqs):
            np.cumsum([len(s) for s in seqs], out=offs[1:])
        # The reference upper-cases what it reads from a FASTA file (fasta_reader.py:55,62) but takes hand-assigned
        # ``seqs`` and the ``seq`` of occurrences(row, seq) as they are: a lower-case letter is then simply not in
        # the k-mer map (kmer_counts.py:146-147).  The packer folds case, so lower-case letters are made invalid here.
        joined = "".join(seqs).encode("latin-1", "replace").translate(_LOWER_TO_INVALID)
        letters = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(1, dtype=np.uint8)
        out = ctypes.c_void_p()
        rc = lib.skr_pack_sequences(ctypes.c_void_p(letters.ctypes.data), ctypes.c_void_p(offs.ctypes.data),
                                    len(seqs), ctypes.c_void_p(lut.ctypes.data), nthreads, int(pinned),
                                    ctypes.byref(out))
        _lib.check(rc)
        obj = cls(out, None)
        obj._seq_list = list(seqs)
        return obj

    def wait(self):
        """Block until a background packer (from_file(..., background=True)) has filled codes and mask."""
        if getattr(self, "_pending", False) and self._h is not None:
            _lib.check(self._lib.skr_packed_wait(self._h))
            self._pending = False
        return self

    def close(self):
        if self._h is not None:
            self._lib.skr_packed_free(self._h)  # waits for a background packer first
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host views -----------------------------------------------------------------------------
    def _view(self, getter, count, dtype):
        addr = getter(self._h)
        if not addr or count == 0:
            return np.zeros(0, dtype=dtype)
        buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype)

    @property
    def lengths(self):
        return self._view(self._lib.skr_packed_lengths, self.m, np.uint32)

    @property
    def block_offsets(self):
        return self._view(self._lib.skr_packed_block_offsets, self.m + 1, np.uint64)

    @property
    def codes(self):
        self.wait()
        return self._view(self._lib.skr_packed_codes, self.nblocks * 4, np.uint32)

    @property
    def mask(self):
        self.wait()
        return self._view(self._lib.skr_packed_mask, self.nblocks * 2, np.uint32)

    def header(self, i):
        spans = self._view(self._lib.skr_packed_header_spans, 2 * self.m, np.uint64)
        off, n = int(spans[2 * i]), int(spans[2 * i + 1])
        return bytes(self._text[off:off + n]).decode("utf-8", "replace")

    def headers(self):
        spans = self._view(self._lib.skr_packed_header_spans, 2 * self.m, np.uint64)
        text = self._text
        return [bytes(text[int(spans[2 * i]):int(spans[2 * i]) + int(spans[2 * i + 1])]).decode("utf-8", "replace")
                for i in range(self.m)]

    def sequence(self, i):
        """Record i as the reference's Reader returns it: lines stripped, joined, upper-cased."""
        if self._text is None:
            return self._seq_list[i]
        spans = self._view(self._lib.skr_packed_body_spans, 2 * self.m, np.uint64)
        off, n = int(spans[2 * i]), int(spans[2 * i + 1])
        body = bytes(self._text[off:off + n]).decode("utf-8", "replace")
        # text-mode line ends (fasta_reader.py:44): \r\n, \r and \n
        lines = body.replace("\r\n", "\n").replace("\r", "\n").split("\n")
        return "".join(line.strip() for line in lines).upper()


class LazySeqs:
    """Read-only list-like view of the records of a PackedFasta (``BasicCounter.seqs``)."""

    def __init__(self, packed):
        self.packed = packed

    def __len__(self):
        return self.packed.m

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.packed.sequence(j) for j in range(*i.indices(self.packed.m))]
        if i < 0:
            i += self.packed.m
        if not 0 <= i < self.packed.m:
            raise IndexError("list index out of range")
        return self.packed.sequence(i)

    def __iter__(self):
        for i in range(self.packed.m):
            yield self.packed.sequence(i)

    def __eq__(self, other):
        return list(self) == list(other)

    def __repr__(self):
        return "LazySeqs(%d records)" % self.packed.m


class Reader:
    """Same interface as the reference's ``Reader`` (seekr/fasta_reader.py:9-109)."""

    def __init__(self, infasta=None, outfasta=None, names=None):
        self.infasta = infasta
        self.outfasta = outfasta
        self.names = names
        self._data = None
        self._packed = None  # parsed but not yet decoded into ``data``

    # ``data`` is filled as a side effect of get_lines / get_headers / get_seqs in the reference
    # (fasta_reader.py:65-78).  get_headers() of a 100 MB file should not decode every sequence into a Python
    # string (300 ms for 30 000 transcripts, against 5 ms for the headers), so the list is built on first use.
    @property
    def data(self):
        if self._data is None and self._packed is not None:
            self._data = self._lines_of(self._packed)
            self._packed = None
        return self._data

    @data.setter
    def data(self, value):
        self._data = value
        self._packed = None

    @staticmethod
    def _lines_of(packed):
        if packed.m == 0:
            return [""]  # what the reference produces for an empty file (fasta_reader.py:62)
        data = []
        headers = packed.headers()
        for i in range(packed.m):
            data.append(headers[i])
            data.append(packed.sequence(i))
        return data

    def _parse(self):
        # any 4-letter alphabet will do: only the record structure is used here
        return PackedFasta.from_file(self.infasta)

    def get_lines(self):
        """[header, SEQ, header, SEQ, ...] (fasta_reader.py:65-68)."""
        self._data = None
        self._packed = self._parse()
        return self.data

    def get_seqs(self):
        return self.get_lines()[1::2]

    def get_headers(self):
        """Header lines (with '>'); ``data`` is decoded only if somebody reads it afterwards."""
        packed = self._parse()
        self._data = None
        self._packed = packed
        if packed.m == 0:
            return self.data[::2]
        return packed.headers()

    def get_data(self, tuples_only=False):
        clean = self.get_lines()
        headers, seqs = clean[::2], clean[1::2]
        tuples = zip(headers, seqs)
        if tuples_only:
            return tuples
        return tuples, headers, seqs

    def supply_basic_header(self):
        """Convert header lines to GENCODE format with only common name and length (fasta_reader.py:88-102)."""
        new_fasta = []
        if self.names is None:
            self.names = iter(self.get_headers())
        for i, line in enumerate(self.data):
            if line[0] == ">":
                name = next(self.names).strip(">")
                length = len(self.data[i + 1])
                new_fasta.append(">||||{}||{}|".format(name, length))
            else:
                new_fasta.append(line)
        return new_fasta

    def save(self):
        with open(self.outfasta, "w") as outfasta:
            for line in self.data:
                outfasta.write(line + "\n")
