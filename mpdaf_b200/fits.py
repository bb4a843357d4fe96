"""Minimal FITS cube I/O for the combine path (no astropy in this image).

Covers what the path touches on either side of the kernels:
  * reading the headers / shape of the MUSE cubes a `CubeList` is built from
    (reference: DataArray.__init__ reads headers only, lib/mpdaf/obj/data.py:292-356);
  * writing a cube as PRIMARY + DATA (+ STAT) image HDUs, float32 big-endian with
    NaN for masked voxels (reference: savemask='nan', data.py:1067-1069,1116-1118);
  * memory-mapping an image HDU as a NumPy array.
Only BITPIX -32/-64/32 images and simple header cards are supported.
"""
import numpy as np

BLOCK = 2880
CARD = 80


def _format_value(value):
    if isinstance(value, bool):
        return ("T" if value else "F").rjust(20)
    if isinstance(value, (int, np.integer)):
        return str(int(value)).rjust(20)
    if isinstance(value, (float, np.floating)):
        text = repr(float(value)).upper()
        if "E" not in text and "." not in text and "N" not in text:
            text += "."
        return text.rjust(20)
    text = str(value).replace("'", "''")
    return ("'" + text.ljust(8) + "'").ljust(20)


def make_card(key, value=None, comment=""):
    key = key.upper()
    if key in ("COMMENT", "HISTORY", "END") or value is None:
        card = key.ljust(8) + (("  " + str(comment)) if comment else "")
        return card[:CARD].ljust(CARD)
    if key.startswith("HIERARCH "):
        card = key + " = " + _format_value(value).strip()
    else:
        card = key[:8].ljust(8) + "= " + _format_value(value)
    if comment:
        card += " / " + str(comment)
    return card[:CARD].ljust(CARD)


def _header_bytes(cards):
    text = "".join(cards) + "END".ljust(CARD)
    pad = (-len(text)) % BLOCK
    return (text + " " * pad).encode("ascii")


def _image_cards(first, arr, extname, extra):
    cards = [first, make_card("BITPIX", {4: -32, 8: -64}[arr.dtype.itemsize] if arr.dtype.kind == "f" else 32),
             make_card("NAXIS", arr.ndim)]
    for k, n in enumerate(reversed(arr.shape)):
        cards.append(make_card(f"NAXIS{k + 1}", n))
    cards += [make_card("PCOUNT", 0), make_card("GCOUNT", 1), make_card("EXTNAME", extname)]
    for key, val in (extra or {}).items():
        if isinstance(val, tuple):
            cards.append(make_card(key, val[0], val[1]))
        else:
            cards.append(make_card(key, val))
    return cards


def write_cube(path, data, var=None, primary_header=None, data_header=None, dtype=np.float32,
               comments=()):
    """Write PRIMARY + DATA [+ STAT] image HDUs. `data`/`var` are (nz, ny, nx) arrays."""
    prim = [make_card("SIMPLE", True, "conforms to FITS standard"), make_card("BITPIX", 8),
            make_card("NAXIS", 0), make_card("EXTEND", True)]
    for key, val in (primary_header or {}).items():
        if isinstance(val, tuple):
            prim.append(make_card(key, val[0], val[1]))
        else:
            prim.append(make_card(key, val))
    for c in comments:
        prim.append(make_card("COMMENT", None, c))
    with open(path, "wb") as f:
        f.write(_header_bytes(prim))
        for name, arr in (("DATA", data), ("STAT", var)):
            if arr is None:
                continue
            arr = np.asarray(arr)
            if arr.dtype.kind == "f":
                be = np.ascontiguousarray(arr, dtype=np.dtype(dtype).newbyteorder(">"))
            else:
                be = np.ascontiguousarray(arr, dtype=">i4")
            cards = _image_cards(make_card("XTENSION", "IMAGE", "Image extension"), be, name,
                                 data_header if name == "DATA" else None)
            f.write(_header_bytes(cards))
            raw = be.tobytes()
            f.write(raw)
            f.write(b"\0" * ((-len(raw)) % BLOCK))


def _parse_card(card):
    key = card[:8].strip()
    if card.startswith("HIERARCH"):
        left, _, right = card.partition("=")
        key, rest = left.strip(), right
    elif card[8:10] == "= ":
        rest = card[10:]
    else:
        return key, card[8:].strip(), True
    rest = rest.strip()
    if rest.startswith("'"):
        out, i = [], 1
        while i < len(rest):
            if rest[i] == "'":
                if i + 1 < len(rest) and rest[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(rest[i])
            i += 1
        return key, "".join(out).rstrip(), False
    val = rest.split("/")[0].strip()
    if val == "T":
        return key, True, False
    if val == "F":
        return key, False, False
    try:
        return key, int(val), False
    except ValueError:
        try:
            return key, float(val), False
        except ValueError:
            return key, val, False


def read_hdus(path):
    """Return [(header dict, data_offset, nbytes)] for every HDU (headers only)."""
    hdus = []
    with open(path, "rb") as f:
        off = 0
        while True:
            header, comments, done = {}, [], False
            while not done:
                blk = f.read(BLOCK)
                if len(blk) < BLOCK:
                    return hdus
                off += BLOCK
                text = blk.decode("ascii", "replace")
                for c in range(0, BLOCK, CARD):
                    card = text[c:c + CARD]
                    if card.startswith("END") and card[3:].strip() == "":
                        done = True
                        break
                    if not card.strip():
                        continue
                    key, val, commentary = _parse_card(card)
                    if commentary:
                        comments.append((key, val))
                    else:
                        header[key] = val
            header["_COMMENTS"] = comments
            naxis = int(header.get("NAXIS", 0))
            nelem = 1 if naxis > 0 else 0
            for k in range(naxis):
                nelem *= int(header[f"NAXIS{k + 1}"])
            nbytes = abs(int(header.get("BITPIX", 8))) // 8 * int(header.get("GCOUNT", 1)) * (
                int(header.get("PCOUNT", 0)) + nelem)
            hdus.append((header, off, nbytes))
            off += (nbytes + BLOCK - 1) // BLOCK * BLOCK
            f.seek(off)


def find_image(path, extname):
    for header, off, nbytes in read_hdus(path):
        if str(header.get("EXTNAME", "")).upper() == extname.upper():
            return header, off, nbytes
    raise KeyError(f"{path}: no HDU named {extname}")


def read_image(path, extname="DATA"):
    """Memory-map image HDU `extname` as a big-endian NumPy array of shape (nz, ny, nx)."""
    header, off, _ = find_image(path, extname)
    dtype = {-32: ">f4", -64: ">f8", 32: ">i4"}[int(header["BITPIX"])]
    shape = tuple(int(header[f"NAXIS{k}"]) for k in range(int(header["NAXIS"]), 0, -1))
    return np.memmap(path, dtype=dtype, mode="r", offset=off, shape=shape)
