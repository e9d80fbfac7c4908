"""Multi-ARFCN receive chain of radioInterfaceMulti (radioInterfaceMulti.cpp:237-362) on device streams:

    wideband block stream -> Channelizer(m, block_len) -> per channel Resampler(p, q) -> 625-sample slots

The reference runs it block by block (one Channelizer::rotate of block_len * m samples, then m Resampler::rotate
calls of block_len samples each, the 16 samples of resampler history being the tail of the channel's previous
block).  Here any number of blocks go through per call: the channelizer writes each channel row behind 16 samples
of carried history, and every channel is one long resampler stream (identical to the block-by-block calls because
q * out_per_block / p == block_len, Resampler.cpp:131-150).

Time-block sharding: a rank that starts in the middle of the stream calls prime() with the
32 wideband rows in front of its first block - re-read from the source - instead of carrying state from a previous
call: rows 0..15 become the channelizer's history (Channelizer.cpp:87-88), rows 16..31 are channelized into the
resampler's history.  Its outputs are then bit-identical to the same blocks of an unsharded run.
"""
import torch

from . import Channelizer, Resampler

HALO_ROWS = 32  # wideband rows (of m samples) a mid-stream start re-reads: 16 channelizer taps + 16 resampler taps


class WidebandRx:
    def __init__(self, trx, m=64, block_len=192, p=65, q=48, filt_len=16):
        assert block_len % q == 0, "radioInterfaceMulti sizes its blocks as a multiple of the resampler period"
        self.trx, self.m, self.block_len, self.p, self.q, self.hist = trx, m, block_len, p, q, filt_len
        self.out_per_block = block_len // q * p
        self.ch = Channelizer(trx, m, block_len, filt_len)
        self.rs = Resampler(trx, p, q, filt_len)
        self.tail = torch.zeros((m, filt_len, 2), dtype=torch.float32, device=trx.device)
        self.buf = None

    def reset(self):
        self.ch.reset()
        self.tail.zero_()

    def prime(self, halo):
        """halo: the HALO_ROWS wideband rows [HALO_ROWS * m, 2] that precede the first block of the next rotate()."""
        assert halo.shape[0] == HALO_ROWS * self.m
        h = self.hist
        self.ch.prime(halo[: h * self.m])
        self.ch.rotate_into(halo[h * self.m:], self.tail, 0)

    def rotate(self, wide, out=None):
        """wide [n_blocks * block_len * m, 2] -> [m, n_blocks * out_per_block, 2] (4-sps channel streams)."""
        m, h = self.m, self.hist
        T = wide.shape[0] // m
        nblk = T // self.block_len
        assert T == nblk * self.block_len
        if self.buf is None or self.buf.shape[1] != h + T:
            self.buf = torch.empty((m, h + T, 2), dtype=torch.float32, device=wide.device)
        buf = self.buf
        buf[:, :h] = self.tail
        self.ch.rotate_into(wide, buf, h)
        L = nblk * self.out_per_block
        if out is None:
            out = torch.empty((m, L, 2), dtype=torch.float32, device=wide.device)
        self.rs.rotate_streams(buf, h, T, buf.stride(0) // 2, m, out, L)
        self.tail.copy_(buf[:, T:T + h])
        return out

    @staticmethod
    def slots(streams, first=0, n_slots=None, slot_len=625):
        """[m, L, 2] channel streams -> [m * n_slots, slot_len, 2] burst rows (a view when the slots tile a row exactly,
        RadioInterface::pullBuffer slices its ring buffer the same way, radioInterface.cpp:257-258)."""
        m, L = streams.shape[0], streams.shape[1]
        if n_slots is None:
            n_slots = (L - first) // slot_len
        s = streams[:, first:first + n_slots * slot_len]
        return s.reshape(m * n_slots, slot_len, 2) if s.is_contiguous() else s.contiguous().reshape(m * n_slots, slot_len, 2)
