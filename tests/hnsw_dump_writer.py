"""Test fixture generator: writes a graph dump in hnsw_rs's on-disk format
(patches/hnsw_rs/src/hnswio.rs: Description::dump :878-918, PointIndexation::dump :1303-1340,
dump_point :1063-1096).  Test infrastructure only -- the product only reads this format."""
import struct

MAGICPOINT, MAGICDESCR_3, MAGICDESCR_4, MAGICLAYER = 0x000A678F, 0x002A6771, 0x002A6779, 0x000A676F
NB_LAYER_MAX = 16


def write_graph_dump(path, points, *, full=True, version=4, max_nb_connection=16, ef=200, dimension=8,
                     distname="DistCosine", t_name="f32", level_scale=1.0, entry=None):
    """points: list of dicts {origin_id, layer, rank, neighbours: [list of (d_id, distance, layer, rank)] x 16},
    grouped by layer (layer block b holds the points whose own top layer is b, rank = index in block)."""
    blocks = {}
    for p in points:
        blocks.setdefault(p["layer"], []).append(p)
    nb_blocks = (max(blocks) + 1) if blocks else 0
    out = bytearray()
    out += struct.pack("=I", MAGICDESCR_4 if version == 4 else MAGICDESCR_3)
    out += struct.pack("=BB", 1 if full else 0, max_nb_connection)
    if version == 4:
        out += struct.pack("=d", level_scale)
    out += struct.pack("=B", NB_LAYER_MAX)
    out += struct.pack("=QQQ", ef, len(points), dimension)
    for name in (distname, t_name):
        b = name.encode()
        out += struct.pack("=Q", len(b)) + b
    out += struct.pack("=B", nb_blocks)
    for b in range(nb_blocks):
        pts = blocks.get(b, [])
        out += struct.pack("=IQ", MAGICLAYER, len(pts))
        for j, p in enumerate(pts):
            assert p["rank"] == j
            out += struct.pack("=IQ", MAGICPOINT, p["origin_id"])
            if full:
                out += struct.pack("=Bi", p["layer"], p["rank"])
            assert len(p["neighbours"]) == NB_LAYER_MAX
            for nl in p["neighbours"]:
                out += struct.pack("=Q", len(nl))
                for (d_id, dist, l, r) in nl:
                    out += struct.pack("=Q", d_id)
                    if full:
                        out += struct.pack("=Bi", l, r)
                    out += struct.pack("=f", dist)
    ep = entry if entry is not None else (points[-1] if points else {"origin_id": 0, "layer": 0, "rank": 0})
    out += struct.pack("=Q", ep["origin_id"])
    if full:
        out += struct.pack("=Bi", ep["layer"], ep["rank"])
    with open(path, "wb") as f:
        f.write(out)
