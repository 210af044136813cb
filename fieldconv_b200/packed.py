"""Host-side mirror of the packed operand format ("PK", fieldconv_b200/csrc/common.cuh) written by the packing
aggregation kernels and bulk-copied by the 2xFP16 contraction kernels.  Used by tests and debugging only: the product
path never decodes a PK buffer.

A real matrix [rows x cols] (cols % 64 == 0) is stored as [row tile of 128][chunk of 64 columns][plane hi, lo] blocks of
16 KB; inside a block row r starts at r*128 bytes and its 16-byte unit u (8 fp16) sits at unit u ^ (r & 7) — the
shared-memory image of a SWIZZLE_128B K-major tile.  hi = fp16(s*v), lo = fp16(s*v - hi), s a power of two that puts
the operand's a-priori bound in [2^14, 2^15)."""
import struct

import torch

ROWS, COLS = 128, 64
PLANE_BYTES = ROWS * COLS * 2
BLOCK_BYTES = 2 * PLANE_BYTES


def padded_rows(rows):
    return (rows + ROWS - 1) // ROWS * ROWS


def pk_bytes(rows, cols):
    assert cols % COLS == 0
    return padded_rows(rows) // ROWS * (cols // COLS) * BLOCK_BYTES


def byte_offset(row, col, cols, plane=0):
    """Byte address of element (row, col) of plane 0 (hi) / 1 (lo) — the arithmetic of store_ring_packed (aggregate.cu)."""
    tile, r = divmod(row, ROWS)
    chunk, k = divmod(col, COLS)
    return ((tile * (cols // COLS) + chunk) * 2 + plane) * PLANE_BYTES + r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2


def scale_of(bound):
    """The power-of-two operand scale the kernels derive from the bound (scale_field in common.cuh)."""
    bits = struct.unpack("<I", struct.pack("<f", float(bound)))[0]
    e = (bits >> 23) & 0xFF
    if e == 255:
        return 1.0
    f = min(max(268 - e, 1), 253)
    return 2.0 ** (f - 127)


def unpack(buf, rows, cols, bound):
    """Decode a PK buffer (any tensor whose storage holds pk_bytes(rows, cols) bytes) -> float32 [padded_rows, cols]
    = (hi + lo) / scale, rows >= `rows` being the zero tail of the last tile."""
    assert cols % COLS == 0
    tiles, chunks = padded_rows(rows) // ROWS, cols // COLS
    if buf.is_complex():
        buf = torch.view_as_real(buf)
    h = buf.contiguous().view(-1).view(torch.float16)[:tiles * chunks * 2 * ROWS * COLS]
    h = h.reshape(tiles, chunks, 2, ROWS, 8, 8)                       # [tile][chunk][plane][row][stored unit][element]
    r = torch.arange(ROWS, device=h.device).reshape(ROWS, 1)
    u = torch.arange(8, device=h.device).reshape(1, 8)
    src = (u ^ (r & 7)).reshape(1, 1, 1, ROWS, 8, 1).expand_as(h)      # logical unit u of row r is stored at u ^ (r & 7)
    h = torch.gather(h, 4, src)
    v = (h[:, :, 0].float() + h[:, :, 1].float()) / scale_of(bound)   # [tile][chunk][row][unit][element]
    return v.permute(0, 2, 1, 3, 4).reshape(tiles * ROWS, cols)
