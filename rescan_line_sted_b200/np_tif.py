"""Minimal uncompressed-TIFF stack I/O (host-side file format helper).

Same call surface as the reference's `figure_generation/np_tif.py`
(`array_to_tif` ref np_tif.py:114, `tif_to_array` ref np_tif.py:13) for the
uses `line_sted_tools` makes of it: ImageJ-readable stacks of 8/16/32-bit
integer or 32/64-bit float images, one strip per image, float64 narrowed to
float32 on disk by default.  Written from the TIFF 6.0 layout with `struct`;
not on the GPU hot path.
"""
import struct

import numpy as np

_TYPE_SIZES = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8,
               11: 4, 12: 8}
_SAMPLE_FORMAT = {'u': 1, 'i': 2, 'f': 3}
_NARROW = {np.dtype('float64'): np.dtype('float32'),
           np.dtype('int64'): np.dtype('int32'),
           np.dtype('uint64'): np.dtype('uint32')}


def _ifd_bytes(width, height, bits, sample_format, desc_offset, desc_len,
               strip_offset, strip_bytes, next_ifd):
    entries = [  # (tag, type, count, value), ascending tags
        (254, 4, 1, 0),                # NewSubfileType
        (256, 4, 1, width),            # ImageWidth
        (257, 4, 1, height),           # ImageLength
        (258, 3, 1, bits),             # BitsPerSample
        (262, 3, 1, 1),                # Photometric: BlackIsZero
        (270, 2, desc_len, desc_offset),  # ImageDescription
        (273, 4, 1, strip_offset),     # StripOffsets
        (277, 3, 1, 1),                # SamplesPerPixel
        (278, 3, 1, height),           # RowsPerStrip
        (279, 4, 1, strip_bytes),      # StripByteCounts
        (339, 3, 1, sample_format),    # SampleFormat
    ]
    out = struct.pack('<H', len(entries))
    for tag, typ, count, value in entries:
        out += struct.pack('<HHII', tag, typ, count, value)
    return out + struct.pack('<I', next_ifd)


def array_to_tif(x, filename, slices=None, channels=None, verbose=False,
                 coerce_64bit_to_32bit=True):
    """Write a 1-, 2- or 3-D array as a little-endian TIFF stack laid out the
    way ImageJ does: header, first IFD, description, pixel data, other IFDs."""
    x = np.asarray(x)
    if x.ndim == 1:
        x = x.reshape((1, 1) + x.shape)
    elif x.ndim == 2:
        x = x.reshape((1,) + x.shape)
    assert x.ndim == 3
    dtype = x.dtype
    if coerce_64bit_to_32bit and dtype in _NARROW:
        dtype = _NARROW[dtype]
    dtype = np.dtype(dtype).newbyteorder('<')
    if dtype.kind not in _SAMPLE_FORMAT:
        raise UserWarning('array_to_tif: unsupported dtype %s' % x.dtype)
    n, h, w = x.shape
    bits = dtype.itemsize * 8
    if slices is not None and channels is not None:
        assert slices * channels == n
        desc = ('ImageJ=1.48e\nimages=%i\nchannels=%i\nslices=%i\n'
                'hyperstack=true\nmode=grayscale\n' % (n, channels, slices))
    else:
        desc = 'ImageJ=1.48e\nimages=%i\nslices=%i\n' % (n, n)
    desc += 'loop=false\nmin=%0.3f\nmax=%0.3f\n\x00' % (x.min(), x.max())
    desc = desc.encode('ascii')
    ifd_len = len(_ifd_bytes(0, 0, 0, 0, 0, 0, 0, 0, 0))
    desc_offset = 8 + ifd_len
    data_offset = desc_offset + len(desc)
    strip_bytes = h * w * dtype.itemsize
    tail_offset = data_offset + n * strip_bytes
    if verbose:
        print('Writing', filename, x.shape, dtype)
    with open(filename, 'wb') as f:
        f.write(b'II*\x00' + struct.pack('<I', 8))
        f.write(_ifd_bytes(w, h, bits, _SAMPLE_FORMAT[dtype.kind], desc_offset,
                           len(desc), data_offset, strip_bytes,
                           tail_offset if n > 1 else 0))
        f.write(desc)
        for z in range(n):  # one image at a time keeps the conversion small
            f.write(np.ascontiguousarray(x[z], dtype=dtype).tobytes())
        for z in range(1, n):
            nxt = tail_offset + z * ifd_len if z < n - 1 else 0
            f.write(_ifd_bytes(w, h, bits, _SAMPLE_FORMAT[dtype.kind],
                               desc_offset, len(desc),
                               data_offset + z * strip_bytes, strip_bytes,
                               nxt))
    return None


def _read_ifd(buf, offset, bo):
    (count,) = struct.unpack_from(bo + 'H', buf, offset)
    tags = {}
    for i in range(count):
        tag, typ, cnt = struct.unpack_from(bo + 'HHI', buf, offset + 2 + 12 * i)
        size = _TYPE_SIZES.get(typ, 1) * cnt
        pos = offset + 2 + 12 * i + 8
        if size > 4:
            (pos,) = struct.unpack_from(bo + 'I', buf, pos)
        fmt = {1: 'B', 2: 'c', 3: 'H', 4: 'I', 6: 'b', 8: 'h', 9: 'i',
               11: 'f', 12: 'd'}.get(typ)
        if fmt is None:  # rationals etc.: not needed here
            continue
        vals = struct.unpack_from(bo + fmt * cnt, buf, pos)
        tags[tag] = vals
    (nxt,) = struct.unpack_from(bo + 'I', buf, offset + 2 + 12 * count)
    return tags, nxt


def tif_to_array(filename, image_descriptor=None, verbose=False):
    """Read an uncompressed greyscale TIFF stack into an (n, h, w) array."""
    with open(filename, 'rb') as f:
        buf = f.read()
    if buf[:2] == b'II':
        bo = '<'
    elif buf[:2] == b'MM':
        bo = '>'
    else:
        raise UserWarning('Not a TIFF file: %s' % filename)
    magic, offset = struct.unpack_from(bo + 'HI', buf, 2)
    if magic != 42:
        raise UserWarning('Not a TIFF file: %s' % filename)
    images = []
    while offset:
        tags, offset = _read_ifd(buf, offset, bo)
        if tags.get(259, (1,))[0] != 1:
            raise UserWarning('Compressed TIFFs are not supported')
        if tags.get(277, (1,))[0] != 1:
            raise UserWarning('Only one sample per pixel is supported')
        w, h = tags[256][0], tags[257][0]
        bits = tags.get(258, (1,))[0]
        kind = {1: 'u', 2: 'i', 3: 'f'}[tags.get(339, (1,))[0]]
        dtype = np.dtype('%s%s%d' % (bo, kind, bits // 8))
        offsets = tags[273]
        counts = tags.get(279, (w * h * bits // 8,))
        raw = b''.join(buf[o:o + c] for o, c in zip(offsets, counts))
        images.append(np.frombuffer(raw, dtype=dtype, count=w * h)
                      .reshape(h, w))
        if verbose:
            print('Read IFD', len(images), (h, w), dtype)
    if not images:
        raise UserWarning('No images in %s' % filename)
    out = np.stack(images, axis=0)
    return out.astype(out.dtype.newbyteorder('='))
