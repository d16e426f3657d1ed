"""Readers/writers for the reference's dump formats.

`<o>.ec.bin`            write_All_reads / load_All_reads (Process_Read.cpp:69-125,127-)
`<o>.ovlp.{source,reverse}.bin`  write_ma_hit_ts / load_ma_hit_ts (Overlaps.cpp:23442,23328)
`<o>.ovlp.paf`          Output_PAF (Assembly.cpp:1673-1717)

Host-side format code (numpy); no compute happens here.
"""
from __future__ import annotations

import dataclasses
import numpy as np

# on-disk ma_hit_t record: 42 bytes, fields written one by one (Overlaps.cpp:23420-23439)
MA_DISK = np.dtype([("qns", "<u8"), ("qe", "<u4"), ("tn", "<u4"), ("ts", "<u4"), ("te", "<u4"),
                    ("el", "u1"), ("no_l_indel", "u1"), ("ml", "<u4"), ("rev", "<u4"),
                    ("bl", "<u4"), ("del", "<u4")])
assert MA_DISK.itemsize == 42

# in-memory mirror used across the C-ABI (include/hifiasm_b200.h: hb_ma_hit_t, 40 bytes)
MA_MEM = np.dtype([("qns", "<u8"), ("qe", "<u4"), ("tn", "<u4"), ("ts", "<u4"), ("te", "<u4"),
                   ("ml", "<u4"), ("rev", "<u4"), ("bl", "<u4"), ("del", "<u4"),
                   ("el", "u1"), ("no_l_indel", "u1"), ("pad", "u1", (6,))])
assert MA_MEM.itemsize == 48


@dataclasses.dataclass
class ReadStore:
    """Flattened All_reads (Process_Read.h:115-146)."""
    length: np.ndarray        # u64[n]
    byte_off: np.ndarray      # u64[n+1]
    packed: np.ndarray        # u8, len/4+1 bytes per read
    n_off: np.ndarray         # u64[n+1]
    n_pos: np.ndarray         # u64
    names: list | None = None
    trio_flag: np.ndarray | None = None
    hom_cov: int = 0
    het_cov: int = 0
    adapter_len: int = 0
    index_size: int = 0
    name_index_size: int = 0
    total_reads_bases: int = 0
    name_blob: bytes = b""
    name_index: np.ndarray | None = None

    @property
    def n(self) -> int:
        return int(self.length.size)

    def decode(self, i: int) -> np.ndarray:
        """recover_UC_Read (Process_Read.cpp:716): codes 0..3, 4 = N."""
        L = int(self.length[i])
        b = self.packed[int(self.byte_off[i]):int(self.byte_off[i + 1])]
        q = np.stack([(b >> 6) & 3, (b >> 4) & 3, (b >> 2) & 3, b & 3], axis=1).reshape(-1)[:L].astype(np.uint8)
        for p in self.n_pos[int(self.n_off[i]):int(self.n_off[i + 1])]:
            q[int(p)] = 4
        return q


def load_ec_bin(path: str) -> ReadStore:
    buf = np.fromfile(path, dtype=np.uint8)
    o = 0

    def take(dt, cnt=1):
        nonlocal o
        a = np.frombuffer(buf, dtype=dt, count=cnt, offset=o)
        o += a.nbytes
        return a
    adapter = int(take("<i4")[0])
    index_size, name_index_size, total_reads, total_bases, total_name_len = (int(x) for x in take("<u8", 5))
    n_off = np.zeros(total_reads + 1, dtype=np.uint64)
    n_pos = []
    for i in range(total_reads):
        nn = int(take("<u8")[0])
        n_off[i + 1] = n_off[i] + nn
        if nn:
            n_pos.append(take("<u8", nn).copy())
    length = take("<u8", total_reads).copy()
    nbytes = length // 4 + 1
    byte_off = np.zeros(total_reads + 1, dtype=np.uint64)
    np.cumsum(nbytes, out=byte_off[1:])
    packed = take("u1", int(byte_off[-1])).copy()
    name_blob = take("u1", total_name_len).tobytes()
    name_index = take("<u8", name_index_size).copy()
    trio = take("u1", total_reads).copy()
    hom, het = (int(x) for x in take("<i4", 2))
    names = [name_blob[int(name_index[i]):int(name_index[i + 1])].decode() for i in range(total_reads)]
    return ReadStore(length=length, byte_off=byte_off, packed=packed, n_off=n_off,
                     n_pos=(np.concatenate(n_pos) if n_pos else np.zeros(0, dtype=np.uint64)),
                     names=names, trio_flag=trio, hom_cov=hom, het_cov=het, adapter_len=adapter,
                     index_size=index_size, name_index_size=name_index_size, total_reads_bases=total_bases,
                     name_blob=name_blob, name_index=name_index)


def canonical_packed(rs: ReadStore) -> np.ndarray:
    """Packed reads with the undefined pad bits/bytes masked (SURVEY.md §8c:
    when len%4==0 the extra byte is never initialised; trailing bits of the
    last used byte are zero-filled by ha_compress_base only for len%4!=0)."""
    out = rs.packed.copy()
    for i in range(rs.n):
        L = int(rs.length[i]); o = int(rs.byte_off[i])
        if L % 4 == 0:
            out[o + L // 4] = 0
        else:
            keep = (0xFF << (2 * (4 - L % 4))) & 0xFF
            out[o + L // 4] &= keep
    return out


def load_ovlp_bin(path: str):
    """-> (records MA_DISK[total], off u64[n+1], is_fully_corrected u8[n], is_abnormal u8[n])"""
    buf = np.fromfile(path, dtype=np.uint8)
    n = int(np.frombuffer(buf, dtype="<i8", count=1)[0])
    o = 8
    off = np.zeros(n + 1, dtype=np.uint64)
    fc = np.zeros(n, dtype=np.uint8); ab = np.zeros(n, dtype=np.uint8)
    chunks = []
    for i in range(n):
        fc[i] = buf[o]; ab[i] = buf[o + 1]
        ln = int(np.frombuffer(buf, dtype="<u4", count=1, offset=o + 2)[0])
        o += 6
        off[i + 1] = off[i] + ln
        if ln:
            chunks.append(np.frombuffer(buf, dtype=MA_DISK, count=ln, offset=o))
            o += ln * 42
    rec = np.concatenate(chunks) if chunks else np.zeros(0, dtype=MA_DISK)
    return rec, off, fc, ab


def write_ovlp_bin(path: str, rec: np.ndarray, off: np.ndarray, fc=None, ab=None) -> None:
    """write_ma_hit_ts (Overlaps.cpp:23442-23465); rec may be MA_DISK or MA_MEM."""
    n = off.size - 1
    if rec.dtype != MA_DISK:
        d = np.zeros(rec.size, dtype=MA_DISK)
        for f in MA_DISK.names:
            d[f] = rec[f]
        rec = d
    with open(path, "wb") as f:
        f.write(np.array([n], dtype="<i8").tobytes())
        for i in range(n):
            s, e = int(off[i]), int(off[i + 1])
            f.write(bytes([int(fc[i]) if fc is not None else 0, int(ab[i]) if ab is not None else 0]))
            f.write(np.array([e - s], dtype="<u4").tobytes())
            f.write(rec[s:e].tobytes())


def disk_to_mem(rec: np.ndarray) -> np.ndarray:
    m = np.zeros(rec.size, dtype=MA_MEM)
    for f in MA_DISK.names:
        m[f] = rec[f]
    return m


def write_paf(path: str, rs: ReadStore, rec: np.ndarray, off: np.ndarray) -> None:
    """Output_PAF (Assembly.cpp:1673-1717): qname qlen qs qe +/- tname tlen ts te ml bl 255"""
    with open(path, "w") as f:
        for i in range(off.size - 1):
            for r in rec[int(off[i]):int(off[i + 1])]:
                qn = int(r["qns"]) >> 32
                f.write("%s\t%d\t%d\t%d\t%c\t%s\t%d\t%d\t%d\t%d\t%d\t255\n" % (
                    rs.names[qn], int(rs.length[qn]), int(r["qns"]) & 0xffffffff, int(r["qe"]),
                    "+-"[int(r["rev"])], rs.names[int(r["tn"])], int(rs.length[int(r["tn"])]),
                    int(r["ts"]), int(r["te"]), int(r["ml"]), int(r["bl"])))


# ---- the native writers of the library (csrc/hostio.cu, include/hifiasm_b200.h: hb_write_*): the product path for the stage's
# on-disk outputs; the numpy writers above stay as test tooling.  No GPU is needed for them (host code of the .so).
def _native():
    from . import engine
    return engine._lib(), engine._p


def _mem(rec: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(rec if rec.dtype == MA_MEM else disk_to_mem(rec))


def _names(rs: ReadStore):
    blob = np.frombuffer(rs.name_blob, dtype=np.uint8) if rs.name_blob else np.zeros(1, np.uint8)
    return np.ascontiguousarray(blob), np.ascontiguousarray(rs.name_index, dtype=np.uint64)


def native_write_paf(path: str, rs: ReadStore, rec: np.ndarray, off: np.ndarray) -> None:
    """hb_write_paf = Output_PAF (Assembly.cpp:1673): rec / off = the final same-haplotype list (R_INF.paf)"""
    import ctypes as C
    L, p = _native(); m = _mem(rec); blob, idx = _names(rs); o = np.ascontiguousarray(off, dtype=np.uint64); ln = np.ascontiguousarray(rs.length, dtype=np.uint64)
    rc = L.hb_write_paf(path.encode(), C.c_uint64(rs.n), p(ln), p(blob), p(idx), p(o), p(m if m.size else np.zeros(1, MA_MEM)))
    if rc:
        raise IOError("hb_write_paf(%s) -> %d" % (path, rc))


def native_write_ec_fa(path: str, rs: ReadStore) -> None:
    """hb_write_ec_fa = Output_corrected_reads (Assembly.cpp:884)"""
    import ctypes as C
    L, p = _native(); blob, idx = _names(rs)
    a = [np.ascontiguousarray(x, dtype=t) for x, t in ((rs.length, np.uint64), (rs.packed, np.uint8), (rs.byte_off, np.uint64), (rs.n_off, np.uint64), (rs.n_pos if rs.n_pos.size else np.zeros(1, np.uint64), np.uint64))]
    rc = L.hb_write_ec_fa(path.encode(), C.c_uint64(rs.n), p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]), p(blob), p(idx))
    if rc:
        raise IOError("hb_write_ec_fa(%s) -> %d" % (path, rc))


def native_write_ovlp_bin(path: str, rec: np.ndarray, off: np.ndarray, fc=None, ab=None) -> None:
    """hb_write_ovlp_bin = write_ma_hit_ts (Overlaps.cpp:23442)"""
    import ctypes as C
    L, p = _native(); m = _mem(rec); o = np.ascontiguousarray(off, dtype=np.uint64); n = o.size - 1
    f = np.ascontiguousarray(fc, dtype=np.uint8) if fc is not None else None; a = np.ascontiguousarray(ab, dtype=np.uint8) if ab is not None else None
    rc = L.hb_write_ovlp_bin(path.encode(), C.c_uint64(n), p(o), p(m if m.size else np.zeros(1, MA_MEM)), p(f), p(a))
    if rc:
        raise IOError("hb_write_ovlp_bin(%s) -> %d" % (path, rc))


def native_write_ec_bin(path: str, rs: ReadStore) -> None:
    """hb_write_ec_bin = write_All_reads (Process_Read.cpp:69)"""
    import ctypes as C
    L, p = _native(); blob, idx = _names(rs)
    a = [np.ascontiguousarray(x, dtype=t) for x, t in ((rs.length, np.uint64), (rs.packed, np.uint8), (rs.byte_off, np.uint64), (rs.n_off, np.uint64), (rs.n_pos if rs.n_pos.size else np.zeros(1, np.uint64), np.uint64))]
    tf = np.ascontiguousarray(rs.trio_flag, dtype=np.uint8) if rs.trio_flag is not None else None
    rc = L.hb_write_ec_bin(path.encode(), C.c_int32(rs.adapter_len), C.c_uint64(rs.index_size), C.c_uint64(rs.name_index_size), C.c_uint64(rs.n), C.c_uint64(rs.total_reads_bases),
                           p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]), p(blob), C.c_uint64(len(rs.name_blob)), p(idx), p(tf), C.c_int32(rs.hom_cov), C.c_int32(rs.het_cov))
    if rc:
        raise IOError("hb_write_ec_bin(%s) -> %d" % (path, rc))


def native_load_reads(paths, adapter_len: int = 0) -> ReadStore:
    """hb_readset_load: FASTA / FASTQ (plain or gzip) -> ReadStore, the reference's ingest (htab.cpp:761-813, ha_compress_base Process_Read.cpp:792)"""
    import ctypes as C
    L, p = _native()

    class RS(C.Structure):
        _fields_ = [(k, C.c_uint64) for k in ("n_reads", "total_bases", "total_name_length", "index_size", "name_index_size")] + \
                   [(k, C.c_void_p) for k in ("read_length", "byte_off", "packed", "n_off", "n_pos", "names", "name_index")]
    if isinstance(paths, str):
        paths = [paths]
    arr = (C.c_char_p * len(paths))(*[q.encode() for q in paths]); rs = RS()
    rc = L.hb_readset_load(arr, C.c_int(len(paths)), C.c_int32(adapter_len), C.byref(rs))
    if rc:
        raise IOError("hb_readset_load(%s) -> %d" % (paths, rc))

    def take(ptr, n, dt):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(int(n) * np.dtype(dt).itemsize,)).view(dt).copy() if n else np.zeros(0, dt)
    n = int(rs.n_reads)
    boff = take(rs.byte_off, n + 1, np.uint64); noff = take(rs.n_off, n + 1, np.uint64)
    out = ReadStore(length=take(rs.read_length, n, np.uint64), byte_off=boff, packed=take(rs.packed, int(boff[-1]), np.uint8), n_off=noff, n_pos=take(rs.n_pos, int(noff[-1]), np.uint64),
                    trio_flag=np.zeros(n, np.uint8), adapter_len=adapter_len, index_size=int(rs.index_size), name_index_size=int(rs.name_index_size), total_reads_bases=int(rs.total_bases),
                    name_blob=take(rs.names, int(rs.total_name_length), np.uint8).tobytes(), name_index=take(rs.name_index, int(rs.name_index_size), np.uint64))
    out.names = [out.name_blob[int(out.name_index[i]):int(out.name_index[i + 1])].decode() for i in range(n)]
    L.hb_readset_free(C.byref(rs))
    return out
