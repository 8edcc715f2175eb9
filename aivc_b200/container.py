"""AIVC bitstream container, in memory (byte-identical to the reference's temp-dir files).

  video : [H_x W_x H_y W_y H_z W_z nb_gop idx_first idx_last : 9 x uint16 BE]   header.py:44-83
          then per GOP  [uint32 BE size][GOP bytes]                             cat_binary_files.py:167-183
  GOP   : [is_LDP u8][nb_chained u16][gop_size u16][round(idx_rate*16) u8]      header.py:129-177
          then per frame in DISPLAY order  [uint32 BE size][frame bytes]        cat_binary_files.py:77-96
  frame : four [uint32 BE size][payload] sections (see entropy.py)
"""
import struct


def gop_header(gop_name, idx_rate=0.):
    toks = gop_name.split('_')
    ldp = 'LDP' in toks
    size = int(toks[-1])
    chained = 0 if ldp else int(toks[0])
    return struct.pack('>BHHB', 1 if ldp else 0, chained, size, int(round(idx_rate * 16)))


def parse_gop_header(b):
    ldp, chained, size, rate = struct.unpack('>BHHB', b[:6])
    return ('LDP_%d' % size) if ldp else ('%d_GOP_%d' % (chained, size)), rate / 16


def pack_gop(gop_name, frames_in_display_order, idx_rate=0.):
    out = gop_header(gop_name, idx_rate)
    for fb in frames_in_display_order:
        out += struct.pack('>I', len(fb)) + fb
    return out


def unpack_gop(b):
    name, rate = parse_gop_header(b)
    pos, frames = 6, []
    while pos < len(b):
        n = struct.unpack('>I', b[pos:pos + 4])[0]
        frames.append(b[pos + 4:pos + 4 + n])
        pos += 4 + n
    return name, rate, frames


def pack_video(dim_x, dim_y, dim_z, gops, idx_first, idx_last):
    out = struct.pack('>9H', dim_x[0], dim_x[1], dim_y[0], dim_y[1], dim_z[0], dim_z[1], len(gops),
                      idx_first, idx_last)
    for g in gops:
        out += struct.pack('>I', len(g)) + g
    return out


def unpack_video(b):
    hx, wx, hy, wy, hz, wz, n, first, last = struct.unpack('>9H', b[:18])
    pos, gops = 18, []
    for _ in range(n):
        size = struct.unpack('>I', b[pos:pos + 4])[0]
        gops.append(b[pos + 4:pos + 4 + size])
        pos += 4 + size
    return {'x': (hx, wx), 'y': (hy, wy), 'z': (hz, wz), 'x_uv': ((hx + 1) // 2, (wx + 1) // 2)}, \
        gops, first, last
