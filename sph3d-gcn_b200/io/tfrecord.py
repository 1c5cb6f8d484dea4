"""TFRecord container and tf.train.Example wire format, read and written without TensorFlow.

The reference stores every S3DIS / ShapeNet / ScanNet block as one tf.train.Example in a TFRecord file
(/root/reference/io/make_tfrecord_s3dis.py:227-242: bytes features xyz_raw, rel_xyz_raw, rgb_raw, seg_label, inner_label,
index_label holding raw little-endian float32 / int32 arrays, int64 features scene_label, scene_idx) and reads them back
with tf.data.TFRecordDataset + tf.parse_single_example (s3dis_seg/train_s3dis.py:145-182).  TensorFlow is a third-party
dependency that is not in this image; both formats are public and restated here:

  TFRecord (tensorflow/core/lib/io/record_writer.cc): per record
      uint64 length | uint32 masked_crc32c(length) | byte data[length] | uint32 masked_crc32c(data)      (little endian)
      masked_crc(c) = ((c >> 15) | (c << 17)) + 0xa282ead8  (mod 2^32),  crc32c = CRC-32/Castagnoli (poly 0x1EDC6F41)
  tf.train.Example (tensorflow/core/example/{example,feature}.proto), protobuf wire format:
      Example  { Features features = 1; }
      Features { map<string, Feature> feature = 1; }          map entry: key = 1 (string), value = 2 (Feature)
      Feature  { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
      BytesList { repeated bytes value = 1; }  FloatList { repeated float value = 1 [packed]; }  Int64List { repeated int64 value = 1 [packed]; }

tests/test_io_cpu.py pins the checksum with the CRC-32C check value (0xE3069283 for b"123456789", RFC 3720 B.4) and the
Example codec against the protobuf runtime (google.protobuf, dynamic descriptors of the three messages).
"""
import struct

import numpy as np

# ---------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
_POLY_REFLECTED = 0x82F63B78


def _make_tables():
    t0 = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (_POLY_REFLECTED if c & 1 else 0)
        t0.append(c)
    tables = [t0]
    for k in range(1, 8):                                   # slicing-by-8 tables
        prev = tables[-1]
        tables.append([(prev[i] >> 8) ^ t0[prev[i] & 0xFF] for i in range(256)])
    return tables


_T = _make_tables()


def crc32c(data, crc=0):
    """CRC-32C of a bytes-like object (slicing-by-8; ~10 MB/s in pure Python, enough for tests and tools)."""
    mv = memoryview(data).cast("B")
    c = crc ^ 0xFFFFFFFF
    n8 = len(mv) // 8 * 8
    t0, t1, t2, t3, t4, t5, t6, t7 = _T
    if n8:
        words = np.frombuffer(mv[:n8], dtype="<u4").tolist()
        for i in range(0, len(words), 2):
            lo = words[i] ^ c
            hi = words[i + 1]
            c = (t7[lo & 0xFF] ^ t6[(lo >> 8) & 0xFF] ^ t5[(lo >> 16) & 0xFF] ^ t4[lo >> 24] ^
                 t3[hi & 0xFF] ^ t2[(hi >> 8) & 0xFF] ^ t1[(hi >> 16) & 0xFF] ^ t0[hi >> 24])
    for b in mv[n8:].tolist():
        c = (c >> 8) ^ t0[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- TFRecord container
class RecordError(ValueError):
    pass


def read_records(path, check_crc=True):
    """Yield the payload of every record of a TFRecord file.  The length checksum is always verified (a corrupt length
    would derail the whole stream); the payload checksum when `check_crc`."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise RecordError("%s: truncated record header" % path)
            length, = struct.unpack("<Q", head[:8])
            if struct.unpack("<I", head[8:])[0] != masked_crc32c(head[:8]):
                raise RecordError("%s: corrupt record length" % path)
            data = f.read(length)
            tail = f.read(4)
            if len(data) < length or len(tail) < 4:
                raise RecordError("%s: truncated record" % path)
            if check_crc and struct.unpack("<I", tail)[0] != masked_crc32c(data):
                raise RecordError("%s: corrupt record payload" % path)
            yield data


def write_records(path, payloads):
    """tf.python_io.TFRecordWriter: one framed record per payload."""
    with open(path, "wb") as f:
        for data in payloads:
            head = struct.pack("<Q", len(data))
            f.write(head)
            f.write(struct.pack("<I", masked_crc32c(head)))
            f.write(data)
            f.write(struct.pack("<I", masked_crc32c(data)))


# ---------------------------------------------------------------------------------------------- protobuf wire format
def _varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise RecordError("malformed varint")


def _fields(buf):
    """(field number, wire type, value) of one message; value = int for varint / fixed, memoryview for length-delimited"""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
            if len(val) < ln:
                raise RecordError("truncated field")
        elif wt == 5:
            val, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise RecordError("unsupported wire type %d" % wt)
        yield num, wt, val


def _parse_feature(buf):
    for num, wt, val in _fields(buf):
        if wt != 2:
            continue
        if num == 1:                                                          # BytesList
            return [bytes(v) for n, w, v in _fields(val) if n == 1 and w == 2]
        if num == 2:                                                          # FloatList (packed or not)
            out = []
            for n, w, v in _fields(val):
                if n == 1 and w == 2:
                    out.extend(np.frombuffer(v, dtype="<f4").tolist())
                elif n == 1 and w == 5:
                    out.append(struct.unpack("<f", struct.pack("<I", v))[0])
            return np.asarray(out, dtype=np.float32)
        if num == 3:                                                          # Int64List (packed or not)
            out = []
            for n, w, v in _fields(val):
                if n == 1 and w == 2:
                    p = 0
                    while p < len(v):
                        x, p = _varint(v, p)
                        out.append(x)
                elif n == 1 and w == 0:
                    out.append(v)
            return np.asarray([x - (1 << 64) if x >= (1 << 63) else x for x in out], dtype=np.int64)
    return []


def parse_example(data):
    """tf.parse_single_example without a schema: {feature name: list of bytes | float32 array | int64 array}."""
    buf = memoryview(data).cast("B")
    features = {}
    for num, wt, val in _fields(buf):
        if num != 1 or wt != 2:
            continue
        for n, w, entry in _fields(val):                                      # Features.feature map entries
            if n != 1 or w != 2:
                continue
            key, feat = None, None
            for en, ew, ev in _fields(entry):
                if en == 1 and ew == 2:
                    key = bytes(ev).decode("utf-8")
                elif en == 2 and ew == 2:
                    feat = ev
            if key is not None:
                features[key] = _parse_feature(feat) if feat is not None else []
    return features


def _enc_varint(x):
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(num, payload):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def make_example(features):
    """tf.train.Example(features=tf.train.Features(feature=...)).SerializeToString() for a dict whose values are
    bytes / list of bytes (BytesList), float arrays (FloatList) or int arrays (Int64List)."""
    entries = b""
    for key in features:                                                      # insertion order, like the python writer
        v = features[key]
        if isinstance(v, (bytes, bytearray)):
            v = [bytes(v)]
        if isinstance(v, (list, tuple)) and v and isinstance(v[0], (bytes, bytearray)):
            feat = _ld(1, b"".join(_ld(1, bytes(b)) for b in v))
        else:
            arr = np.asarray(v)
            if arr.dtype.kind == "f":
                feat = _ld(2, _ld(1, arr.astype("<f4").tobytes()))
            else:
                feat = _ld(3, _ld(1, b"".join(_enc_varint(int(x)) for x in arr.reshape(-1))))
        entries += _ld(1, _ld(1, key.encode("utf-8")) + _ld(2, feat))
    return _ld(1, entries)


# ---------------------------------------------------------------------------------------------- dataset plumbing
def shuffled_records(filelist, buffer_size=10000, rng=None, check_crc=True):
    """tf.data.TFRecordDataset(filelist).shuffle(buffer_size): records pass through a reservoir of `buffer_size`
    entries from which one is drawn uniformly each time a new one arrives."""
    rng = np.random.default_rng() if rng is None else rng
    buf = []
    for path in filelist:
        for rec in read_records(path, check_crc=check_crc):
            buf.append(rec)
            if len(buf) > buffer_size:
                yield buf.pop(int(rng.integers(len(buf))))
    while buf:
        yield buf.pop(int(rng.integers(len(buf))))


def batched(items, batch_size):
    """lists of up to batch_size consecutive items (drop_remainder=False)"""
    batch = []
    for it in items:
        batch.append(it)
        if len(batch) == batch_size:
            yield batch
            batch = []
    if batch:
        yield batch
