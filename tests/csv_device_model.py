"""A Python model of the DEVICE algorithm of kq_csv_scan (csrc/kq_csv.cu) — the same bit-mask formulas per 64-byte block,
the separator array, field bounds and value extraction, with Python integers for the 64-bit masks — so that the
algorithm (block-boundary carries, the shifted-mask test for empty records, quote state by prefix XOR) can be checked
against the oracle on the CPU with generated inputs. It is a model for tests, not a fallback: nothing imports it but tests/."""
M64 = (1 << 64) - 1
BLOCK = 64


OUT, IN, OUTE = 0, 1, 2          # quote state between two bytes: outside / inside a quoted section / outside, right behind the closing quote


def _first_record(text: bytes, delims: bytes, term: int):
    """The first non-empty record [b, e) under rule C2 with `delims` as field delimiters, and its delimiters outside
    quotes: (b, e, positions) or None. Byte-at-a-time (host_first_line / host_first_record of kq_csv.cu)."""
    start, state, fresh, seps = 0, OUT, True, []
    for p in range(len(text) + 1):
        at_end = p == len(text)
        c = term if at_end else text[p]
        if c == 0x22 and not at_end:
            if state == IN:
                state = OUTE
            elif state == OUTE:
                state = IN
            else:
                if fresh:
                    state = IN
                fresh = False
            continue
        if state == OUTE:
            state = OUT
        if state == IN and not at_end:
            continue
        if c == term:
            empty = p == start or (term == 0x0A and p == start + 1 and text[start] == 0x0D)
            if not empty:
                return start, p, seps
            start, fresh, seps = p + 1, True, []
        elif c in delims:
            seps.append(p)
            fresh = True
        elif c > 0x20:
            fresh = False
    return None


def detect(text: bytes):
    """host_detect of kq_csv.cu: (delimiter, terminator). The first record is found with all four candidates acting as
    delimiters; the most frequent of them outside quotes wins (ties: in the order , ; TAB |)."""
    term = 0x0A
    if b"\n" not in text and b"\r" in text:
        term = 0x0D
    delim = ord(",")
    rec = _first_record(text, b",;\t|", term)
    if rec is not None:
        best = 0
        for cand in b",;\t|":
            k = sum(1 for p in rec[2] if text[p] == cand)
            if k > best:
                best, delim = k, cand
    return delim, term


def first_record_fields(text: bytes, delim: int, term: int) -> int:
    """host_first_record: the number of fields of the first non-empty record (the file's column count)."""
    rec = _first_record(text, bytes([delim]), term)
    return 0 if rec is None else len(rec[2]) + 1


def raw_masks(text: bytes, n: int, i: int, delim: int, term: int):
    """One bit per byte of block i: quotes, terminators, CRs, delimiters, blanks (<= 0x20 and neither delimiter nor terminator)."""
    b = i * BLOCK
    blk = text[b:min(b + BLOCK, n)]
    qm = tm = cm = dm = bm = 0
    for j, c in enumerate(blk):
        qm |= (c == 0x22) << j
        tm |= (c == term) << j
        cm |= (c == 0x0D) << j
        dm |= (c == delim) << j
        bm |= (c <= 0x20 and c != term and c != delim) << j
    return len(blk), qm, tm, cm, dm, bm


def opener_candidates(text: bytes, b: int, qm: int, tm: int, dm: int, bm: int, delim: int, term: int):
    """Quotes whose nearest non-blank byte in front of them is a delimiter or terminator (or the start of the text): the
    only quotes that can open a quoted section. Field starts are propagated through blank runs with one addition."""
    p = b - 1
    while p >= 0 and text[p] <= 0x20 and text[p] != term and text[p] != delim:
        p -= 1
    carry = int(p < 0 or text[p] == term or text[p] == delim)
    st = (((dm | tm) << 1) & M64) | carry
    reach = ((((bm + (st & bm)) & M64) ^ bm) | st) & M64
    return qm & reach


def quote_walk(nbytes: int, qm: int, cand: int, s: int):
    """The quote state behind the block for state `s` in front of it, and the structural quotes (those that change
    between inside and outside) — quote_walk of kq_csv.cu."""
    sq, prev, m = 0, -1, qm
    while m:
        j = (m & -m).bit_length() - 1
        m &= m - 1
        if s == OUTE and j != prev + 1:
            s = OUT
        if s == IN:
            s, sq = OUTE, sq | (1 << j)
        elif s == OUTE:
            s, sq = IN, sq | (1 << j)
        elif (cand >> j) & 1:
            s, sq = IN, sq | (1 << j)
        prev = j
    if s == OUTE and prev != nbytes - 1:
        s = OUT
    return s, sq


def block_map(text: bytes, n: int, i: int, delim: int, term: int):
    """k_csv_quote_maps: the block's transition (state behind it for each of the three states in front of it)."""
    nb, qm, tm, cm, dm, bm = raw_masks(text, n, i, delim, term)
    cand = opener_candidates(text, i * BLOCK, qm, tm, dm, bm, delim, term) if qm else 0
    return [quote_walk(nb, qm, cand, s)[0] for s in (OUT, IN, OUTE)]


def block_states(text: bytes, n: int, delim: int, term: int, chunk=4):
    """k_csv_quote_maps -> k_csv_compose_chunks -> k_csv_chunk_states -> k_csv_block_states: the quote state in front of
    every block and behind the text, through per-chunk compositions of the block maps."""
    nblocks = (n + BLOCK - 1) // BLOCK
    maps = [block_map(text, n, i, delim, term) for i in range(nblocks)]
    nchunks = (nblocks + chunk - 1) // chunk
    cmaps = []
    for k in range(nchunks):
        m = [OUT, IN, OUTE]
        for i in range(k * chunk, min((k + 1) * chunk, nblocks)):
            m = [maps[i][x] for x in m]
        cmaps.append(m)
    s, chunk_in = OUT, []
    for k in range(nchunks):
        chunk_in.append(s)
        s = cmaps[k][s]
    final, state_in = s, [0] * nblocks
    for k in range(nchunks):
        s = chunk_in[k]
        for i in range(k * chunk, min((k + 1) * chunk, nblocks)):
            state_in[i] = s
            s = maps[i][s]
    return state_in, final


def block_masks(text: bytes, n: int, i: int, state_in: int, delim: int, term: int):
    """csv_block_masks: (rec, delim) bit masks of block i."""
    b = i * BLOCK
    nb, qm, tm, cm, dm, bm = raw_masks(text, n, i, delim, term)
    inside = 0
    if qm or state_in == IN:
        cand = opener_candidates(text, b, qm, tm, dm, bm, delim, term) if qm else 0
        x = quote_walk(nb, qm, cand, state_in)[1]
        for sh in (1, 2, 4, 8, 16, 32):
            x ^= (x << sh) & M64
        inside = (~x & M64) if state_in == IN else x
    t1 = b == 0 or text[b - 1] == term
    t2 = b <= 1 or text[b - 2] == term
    empty = ((tm << 1) & M64) | int(t1)
    if term == 0x0A:
        prev_c = ((cm << 1) & M64) | int(b > 0 and text[b - 1] == 0x0D)
        prev2_t = ((tm << 2) & M64) | (int(t1) << 1) | int(t2)
        empty |= prev_c & prev2_t
    live = (1 << nb) - 1
    return tm & ~inside & ~empty & live, dm & ~inside & live


def value(text: bytes, first: int, last: int) -> bytes:
    """csv_value: trimmed, unquoted bytes of the raw field [first, last)."""
    while first < last and text[first] <= 0x20:
        first += 1
    while last > first and text[last - 1] <= 0x20:
        last -= 1
    if first < last and text[first] == 0x22:
        p = q = first + 1
        while q < last and not (text[q] == 0x22 and not (q + 1 < last and text[q + 1] == 0x22)):
            q += 2 if text[q] == 0x22 else 1
        q = min(q, last)
        while p < q and text[p] <= 0x20:
            p += 1
        while q > p and text[q - 1] <= 0x20:
            q -= 1
        out, j = bytearray(), p
        while j < q:
            out.append(text[j])
            if text[j] == 0x22 and j + 1 < q and text[j + 1] == 0x22:
                j += 1
            j += 1
        return bytes(out)
    return text[first:last]


def scan_resident(text: bytes, n: int, delim: int, term: int, ncols: int, skip: int, partial: bool):
    """csv_scan_resident: passes 1-6 over text[:n] -> (columns of the ncols file columns, consumed). A whole text
    (partial=False) must balance its quotes; a reader's piece (partial=True) may stop anywhere, `consumed` is the byte
    after its last complete record and a piece without one is an error."""
    nblocks = (n + BLOCK - 1) // BLOCK
    state_in, final = block_states(text, n, delim, term)
    sep, rec_last = [], []
    for i in range(nblocks):
        rec, dl = block_masks(text, n, i, state_in[i], delim, term)
        allm = rec | dl
        while allm:
            j = (allm & -allm).bit_length() - 1
            sep.append(i * BLOCK + j)
            if (rec >> j) & 1:
                rec_last.append(len(sep) - 1)
            allm &= allm - 1
    if not partial and final == IN:
        raise ValueError("CSV text ends inside a quoted field")
    nrec = len(rec_last)
    if partial and nrec == 0:
        raise OverflowError("CSV record longer than the reader's piece")
    consumed = sep[rec_last[-1]] + 1 if partial else n
    skip = skip if nrec else 0
    cols = [[] for _ in range(ncols)]
    for rec in range(skip, nrec):
        k0 = rec_last[rec - 1] + 1 if rec else 0
        k1 = rec_last[rec]
        for c in range(ncols):
            if k0 + c > k1:
                cols[c].append("")
                continue
            a = sep[k0 + c - 1] + 1 if k0 + c else 0
            cols[c].append(value(text, a, sep[k0 + c]).decode("utf-8"))
    return cols, consumed


def scan(text: bytes, has_headers=True):
    """kq_csv_scan's passes 1-6 -> list of columns (lists of str), all file columns."""
    if not text:
        return []
    delim, term = detect(text)
    if text[-1] != term:
        text = text + bytes([term])
    ncols = first_record_fields(text, delim, term)
    if ncols == 0:
        return []
    return scan_resident(text, len(text), delim, term, ncols, 1 if has_headers else 0, False)[0]


def reader(text: bytes, has_headers=True, piece=256):
    """kq_csv_reader_open/next: the text streams through two buffers [reserve R][payload R][terminator]; a piece is cut
    after its last complete record, the unfinished tail is moved in front of the next payload, and the gap down to a
    16-byte boundary is filled with terminators (empty lines). Yields the column lists of every batch that has rows."""
    assert piece % 16 == 0 and piece >= 32
    delim, term = detect(text)
    ncols = first_record_fields(text, delim, term)
    R = piece
    buf = [bytearray(2 * R + 64), bytearray(2 * R + 64)]
    start, length, last = [R, R], [0, 0], [False, False]
    pos = 0

    def upload(slot):
        nonlocal pos
        length[slot] = min(R, len(text) - pos)
        last[slot] = pos + length[slot] == len(text)
        buf[slot][R:R + length[slot]] = text[pos:pos + length[slot]]
        pos += length[slot]

    upload(0)
    cur, first, done = 0, True, False
    while not done:
        nxt = cur ^ 1
        if not last[cur]:
            upload(nxt)
        n = (R - start[cur]) + length[cur]
        if last[cur] and text and text[-1] != term:
            buf[cur][R + length[cur]] = term
            n += 1
        t = bytes(buf[cur][start[cur]:start[cur] + n])
        cols, consumed = scan_resident(t, n, delim, term, ncols, 1 if first and has_headers else 0, not last[cur])
        first = False
        if last[cur]:
            done = True
        else:
            carry = n - consumed
            if carry > R - 16:
                raise OverflowError("CSV record longer than the reader's piece")
            start[nxt] = (R - carry) // 16 * 16
            buf[nxt][R - carry:R] = t[consumed:]
            buf[nxt][start[nxt]:R - carry] = bytes([term]) * (R - carry - start[nxt])
        cur = nxt
        if cols and cols[0]:
            yield cols
