"""Host-side batch loop over the C ABI, in Python (used by tests and bench.py's e2e leg).

Mirrors what the C++ CLI host (host/trimmer.cpp) does around the device path, i.e. the part of
Trim_Single::trim_main / Trim_Paired::trim_main (reference src/trim_single.cpp:239-340,
src/trim_paired.cpp:280-453) that stays on the host: cut the input into batches, carry the
incomplete tail into the next batch (GZReader's last_remainder, src/GZReader.cpp:104-129), append
the output streams in order, add up the counters.
"""
from __future__ import annotations

import ctypes as C

from . import capi


class DataError(Exception):
    """A FASTQ data error the reference reports with exit(1)."""

    def __init__(self, kind, file, record, position, byte, lines):
        super().__init__("data error kind %d in file %d record %d" % (kind, file, record))
        self.kind, self.file, self.record, self.position, self.byte, self.lines = kind, file, record, position, byte, lines


def _patch_eof(buf, end, data_len, file_pos_end):
    """The reference drops the last character of an unterminated final line (GZReader.cpp:81-88):
    overwrite it with the newline the reader would have seen."""
    if end > 0 and file_pos_end == data_len and buf[end - 1] != b"\n":
        buf[end - 1] = b"\n"


def trim_stream(ctx: capi.Context, in0: bytes, in1: bytes = b"", pipelined: bool = False, headroom: int | None = None):
    """Run whole inputs through the context.  Returns dict(out=[bytes]*3, counters=dict, batches=int,
    kernel_ms=float, records=[n0, n1]).  Raises DataError on the first data error."""
    n_in = 2 if ctx.params.mode == capi.MODE_PE_2FILE else 1
    data = [in0, in1]
    outs = [[], [], []]
    totals = dict.fromkeys(capi.Result.COUNTERS, 0)
    records = [0, 0]
    stats = dict(batches=0, kernel_ms=0.0, launches=0, fused_batches=0)

    def absorb(res, slot, base_records):
        if res.error.kind:
            e = res.error
            buf = ctx.in_buffer(slot, e.file)
            lines = [bytes(buf[e.line_off[k]:e.line_off[k] + e.line_len[k]]) for k in range(4)]
            raise DataError(e.kind, e.file, base_records[e.file] + e.record, e.position, e.byte, lines)
        for s in range(3):
            if res.out_bytes[s]:
                outs[s].append(ctx.out_bytes(res, s))
        for k in totals:
            totals[k] += getattr(res, k)
        stats["batches"] += 1
        stats["kernel_ms"] += res.kernel_ms
        stats["launches"] += res.kernel_launches
        stats["fused_batches"] += res.fused
        for i in range(2):
            records[i] += res.records[i]

    if not pipelined or n_in == 2:
        # one slot, synchronous: buffer = carried tail + as many new bytes as fit
        pos = [0, 0]          # file offset of the first byte not yet consumed
        while True:
            ends = [0, 0]
            for i in range(n_in):
                chunk = data[i][pos[i]:pos[i] + ctx.slot_bytes]
                buf = ctx.in_buffer(0, i)
                C.memmove(buf, chunk, len(chunk))
                ends[i] = len(chunk)
                _patch_eof(buf, ends[i], len(data[i]), pos[i] + len(chunk))
            if not any(ends) or (n_in == 2 and not all(ends[:2])):
                break         # (two files: no pair can be formed once either file is exhausted, as host/trimmer.cpp does)
            ctx.submit(0, 0, ends[0], 0, ends[1])
            res = ctx.wait(0)
            absorb(res, 0, list(records))
            if not any(res.consumed[i] for i in range(n_in)):
                # (a full slot of one file facing the other file's last, incomplete record is the end of the pairs)
                if any(ends[i] == ctx.slot_bytes and bytes(ctx.in_buffer(0, i)[:ends[i]]).count(b"\n") < 4 for i in range(n_in)):
                    raise capi.SickleError("a record does not fit in a %d-byte slot" % ctx.slot_bytes)
                break  # only an incomplete record (pair) is left: dropped, as the reference does at EOF
            for i in range(n_in):
                pos[i] += res.consumed[i]
    else:
        # pipelined, single input: bulk of the next slot uploaded early at `headroom`, tail copied in front
        S = ctx.slot_bytes
        H = headroom if headroom is not None else max(16, S // 4)
        bulk_cap = S - H
        nslots = ctx.n_slots
        fpos = 0
        tail = b""
        pending = []  # (slot, start, end)
        slot = 0
        d = data[0]
        done_reading = False
        while True:
            # fill + early-upload the next slot's bulk
            bulk = d[fpos:fpos + bulk_cap]
            buf = ctx.in_buffer(slot, 0)
            C.memmove(C.addressof(buf) + H, bulk, len(bulk))
            fpos += len(bulk)
            if fpos == len(d):
                done_reading = True
            if bulk:
                ctx.upload(slot, 0, H, len(bulk))
            # the previous batch must be finished to know the carried tail
            if pending:
                pslot, pstart, pend = pending.pop(0)
                res = ctx.wait(pslot)
                absorb(res, pslot, list(records))
                pbuf = ctx.in_buffer(pslot, 0)
                tail = bytes(pbuf[pstart + res.consumed[0]:pend])
                if res.consumed[0] == 0 and not bulk:
                    break
            if len(tail) > H:
                raise capi.SickleError("carried tail (%d bytes) exceeds the headroom (%d)" % (len(tail), H))
            if not bulk and not tail:
                break
            start = H - len(tail)
            C.memmove(C.addressof(buf) + start, tail, len(tail))
            end = H + len(bulk)
            _patch_eof(buf, end, len(d), fpos if done_reading else -1)
            ctx.submit(slot, start, end)
            pending.append((slot, start, end))
            slot = (slot + 1) % nslots
            if not bulk:
                # nothing new to read: drain and stop (whatever is left is an incomplete record)
                pslot, pstart, pend = pending.pop(0)
                res = ctx.wait(pslot)
                absorb(res, pslot, list(records))
                break
    return dict(out=[b"".join(o) for o in outs], counters=totals, records=records, **stats)


def reference_batches(data: bytes, batch_len: int, minlines: int):
    """Byte ranges of the batches GZReader::read_lines would form (src/GZReader.cpp:59-132): read lines
    until the sum of line lengths (without '\\n') reaches batch_len, keep a multiple of `minlines`
    lines, carry the rest.  Host logic needed only to reproduce the reference's `-a N` output order,
    which is defined per reference batch."""
    ranges = []
    pos = 0            # next unread byte
    carry_start = 0    # start of carried lines
    carry_lens = []
    n = len(data)
    eof = False
    while not eof:
        remaining = batch_len - sum(carry_lens)
        lens = list(carry_lens)
        ends = []      # end offsets (exclusive, after '\n') of lines in this batch, incl. carried ones
        p = carry_start
        for ln in carry_lens:
            p += ln + 1
            ends.append(p)
        while True:
            if pos >= n:
                eof = True
                break
            nl = data.find(b"\n", pos)
            got = (nl - pos + 1) if nl >= 0 else n - pos
            lens.append(got - 1)
            remaining -= got - 1
            pos += got
            ends.append(pos)
            if remaining <= 0:
                break
        extra = len(lens) % minlines
        keep = len(lens) - extra
        if keep == 0:
            break
        end = ends[keep - 1]
        ranges.append((carry_start, end))
        carry_start = end
        carry_lens = lens[keep:]
    return ranges


def recommended_batch_len(file_size: int, b_mib: int = 512, paired: bool = False) -> int:
    """Trim_Single::recommended_batch_len / Trim_Paired::recommended_batch_len
    (src/trim_single.cpp:194-211, src/trim_paired.cpp:246-263)."""
    mx = (b_mib * 1024 * 1024) & 0xFFFFFFFF
    if paired:
        mx //= 2
    rec = file_size // 8
    return 20 if rec < 20 else (mx if rec > mx else rec)


def trim_stream_reference_order(ctx: capi.Context, in0: bytes, in1: bytes = b"", b_mib: int = 512):
    """Like trim_stream, but batches follow the reference's batch geometry so that
    params.emulate_threads = N reproduces `sickle -a N` byte for byte."""
    mode = ctx.params.mode
    paired = mode != capi.MODE_SE
    bl = recommended_batch_len(len(in0), b_mib, paired)
    minlines = 8 if mode in (capi.MODE_PE_INTER, capi.MODE_PE_INTER_M) else 4
    r0 = reference_batches(in0, bl, minlines)
    r1 = reference_batches(in1, bl, 4) if mode == capi.MODE_PE_2FILE else [(0, 0)] * len(r0)
    outs = [[], [], []]
    totals = dict.fromkeys(capi.Result.COUNTERS, 0)
    stats = dict(batches=0, fused_batches=0)
    for (a0, e0), (a1, e1) in zip(r0, r1):
        for i, (a, e, d) in enumerate(((a0, e0, in0), (a1, e1, in1))):
            if e > a:
                if e - a > ctx.slot_bytes:
                    raise capi.SickleError("reference batch (%d bytes) larger than the slot" % (e - a))
                buf = ctx.in_buffer(0, i)
                C.memmove(buf, d[a:e], e - a)
                _patch_eof(buf, e - a, len(d), e)
        ctx.submit(0, 0, e0 - a0, 0, e1 - a1)
        res = ctx.wait(0)
        if res.error.kind:
            raise DataError(res.error.kind, res.error.file, res.error.record, res.error.position, res.error.byte, [])
        for s in range(3):
            if res.out_bytes[s]:
                outs[s].append(ctx.out_bytes(res, s))
        for k in totals:
            totals[k] += getattr(res, k)
        stats["batches"] += 1
        stats["fused_batches"] += res.fused
    return dict(out=[b"".join(o) for o in outs], counters=totals, **stats)
