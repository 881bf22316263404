"""Host-side mirror of perceive_core::search::Searcher against a SQLite database built by executing the
reference's own migrations (crates/perceive-core/migrations/0000{1,2,3}*.sql, copies under tests/golden/):
the load SQL of search.rs:87-92, rebuild_source (search.rs:58-79) and the hydrate
step of search_vector_and_retrieve (search.rs:195-247)."""
import sqlite3
from pathlib import Path

import numpy as np
import pytest

MIGRATIONS = Path(__file__).parent / "golden" / "reference_migrations"


def apply_reference_schema(conn) -> None:
    """What the reference does when it opens its database (crates/perceive-core/db.rs:93-108): foreign keys
    on, then its three migrations in order — executed VERBATIM from the byte-for-byte copies under
    tests/golden/reference_migrations/ (provenance in the README there), so both loaders meet the
    reference's own tables, column types, foreign keys and model rows, not a restatement of them."""
    conn.execute("PRAGMA foreign_keys = ON")
    for name in ("00001_init.sql", "00002_tags.sql", "00003_model_7.sql"):
        conn.executescript((MIGRATIONS / name).read_text())
    conn.execute("PRAGMA user_version = 3")  # rusqlite_migration records the applied count there


DIM = 384


def make_db(orc, n=600, seed=11):
    """3 sources with interleaved item ids; a few skipped / hidden rows; a second model's rows."""
    import perceive_b200 as pb
    conn = sqlite3.connect(":memory:")
    apply_reference_schema(conn)
    for s in (1, 2, 3):
        conn.execute("INSERT INTO sources (id, name, location, compare_strategy, status) VALUES (?,?,?,?,?)",
                     (s, f"src{s}", "/tmp", "mtime", "ready"))
    vecs = orc.synth_rows(seed, 0, 0, n, DIM)
    rows = {}
    for i in range(n):
        item_id = 1000 + i
        source = 1 + (i % 3)
        skipped = "TooLarge" if i % 97 == 5 else None
        hidden = 1700000000 if i % 89 == 7 else None
        conn.execute("INSERT INTO items (id, source_id, external_id, hash, content, name, skipped, hidden_at) "
                     "VALUES (?,?,?,?,?,?,?,?)", (item_id, source, f"doc{i}", "h", f"text {i}", f"name{i}", skipped, hidden))
        conn.execute("INSERT INTO item_embeddings VALUES (7, 0, ?, 0, ?)", (item_id, pb.serialize_embedding(vecs[i])))
        if i % 50 == 0:  # another model version must not leak into the index
            conn.execute("INSERT INTO item_embeddings VALUES (3, 0, ?, 0, ?)", (item_id, pb.serialize_embedding(-vecs[i])))
        if skipped is None and hidden is None:
            rows[item_id] = (source, vecs[i])
    conn.commit()
    return conn, rows


def test_load_rows_follows_the_reference_sql(pcv_lib, orc):
    """CPU only: the decode half of build_sources (search.rs:87-113)."""
    from perceive_b200 import searcher
    conn, live = make_db(orc)
    rows, ids, srcs, dim = searcher._load_rows(conn, 7, 0, [1, 2, 3])
    assert dim == DIM and rows.shape == (len(live), DIM)
    assert set(ids.tolist()) == set(live)
    for r, i, s in zip(rows, ids, srcs):
        assert live[int(i)][0] == int(s) and np.array_equal(r, live[int(i)][1])
    # unlisted sources are dropped (search.rs:107-112)
    _, ids13, srcs13, _ = searcher._load_rows(conn, 7, 0, [1, 3])
    assert set(srcs13.tolist()) == {1, 3} and set(ids13.tolist()) == {i for i, (s, _) in live.items() if s != 2}
    # another model id sees only its own rows
    _, ids_m3, _, _ = searcher._load_rows(conn, 3, 0, [1, 2, 3])
    assert 0 < len(ids_m3) < len(live)


def test_embedding_codec_roundtrip(pcv_lib, orc):
    """search.rs:281-294: little-endian f32, no header."""
    import perceive_b200 as pb
    v = orc.synth_rows(3, 0, 0, 1, DIM)[0]
    blob = pb.serialize_embedding(v)
    assert blob == v.astype("<f4").tobytes() == orc.encode_embedding(v)
    assert np.array_equal(pb.deserialize_embedding(blob), v)
    with pytest.raises(pb.PcvError):
        pb.deserialize_embedding(blob[:-1])  # the reference panics on a trailing partial chunk


@pytest.mark.gpu
def test_searcher_build_search_retrieve_rebuild(pcv_lib, orc):
    import perceive_b200 as pb
    conn, live = make_db(orc)
    ids = np.array(sorted(live), dtype=np.int64)
    rows = np.stack([live[int(i)][1] for i in ids])
    srcs = np.array([live[int(i)][0] for i in ids], dtype=np.int64)
    q = orc.synth_rows(12, 0, 0, 1, DIM)[0]
    s = pb.Searcher.build(conn, 7, 0)
    try:
        for flt in ([1, 2, 3], [2], [1, 3], []):
            got = s.search_vector(flt, 20, q)
            w_ids, w_scores, _ = orc.search(rows, ids, q, 20, source_ids=srcs, sources=flt, mode=orc.MODE_F32_V1)
            assert [g.id for g in got] == w_ids.tolist(), flt
            assert np.array_equal(np.array([g.score for g in got], dtype=np.float32), w_scores)
        # hide one hit AFTER the build: search_vector still returns it (the reference never
        # reads `hidden`, search.rs:34 vs :157-182), the hydrate query drops it (search.rs:210-212)
        top = s.search_vector([1, 2, 3], 5, q)
        conn.execute("UPDATE items SET hidden_at = 1 WHERE id = ?", (top[1].id,))
        s.hidden.add(top[1].id)
        hyd = s.search_vector_and_retrieve(conn, [1, 2, 3], 5, q)
        assert [it.id for _, it in hyd] == [t.id for t in top if t.id != top[1].id]
        assert all(row["id"] == it.id and row["name"].startswith("name") for row, it in hyd)
        scores = [it.score for _, it in hyd]
        assert scores == sorted(scores)  # ascending distance (search.rs:245)
        # rebuild_source picks up the hide and a new item, other sources untouched (search.rs:58-79)
        new_vec = (q / np.linalg.norm(q)).astype(np.float32)
        conn.execute("INSERT INTO items (id, source_id, external_id, hash, content, name) VALUES (5000, ?, 'new', 'h', 'new', 'name-new')",
                     (live[top[1].id][0],))
        conn.execute("INSERT INTO item_embeddings VALUES (7, 0, 5000, 0, ?)", (pb.serialize_embedding(new_vec),))
        s.rebuild_source(conn, live[top[1].id][0], 7, 0)
        again = s.search_vector([1, 2, 3], 5, q)
        assert again[0].id == 5000 and top[1].id not in [a.id for a in again]
    finally:
        s.close()


@pytest.mark.gpu
def test_searcher_filter_hidden_and_like(pcv_lib, orc):
    """Opt-in `filter_hidden`: `hide` (perceive-cli/cmd/hide.rs:9-17) updates the
    database and inserts into `Searcher.hidden`; the reference search ignores the set, so the hydrate
    query (search.rs:210-212) shortens the result.  With the filter on, k visible items come back.
    Also the `--like ID` query (perceive-cli/cmd/search.rs:64-85) from the resident matrix."""
    import perceive_b200 as pb
    conn, live = make_db(orc)
    ids = np.array(sorted(live), dtype=np.int64)
    rows = np.stack([live[int(i)][1] for i in ids])
    q = orc.synth_rows(12, 0, 0, 1, DIM)[0]
    s = pb.Searcher.build(conn, 7, 0)
    try:
        top = s.search_vector([1, 2, 3], 5, q)
        victims = [top[0].id, top[3].id]
        for v in victims:  # what `perceive hide` does
            conn.execute("UPDATE items SET hidden_at = 1 WHERE id = ?", (v,))
            s.hidden.add(v)
        assert len(s.search_vector_and_retrieve(conn, [1, 2, 3], 5, q)) == 3  # reference behaviour: short result
        s.filter_hidden = True
        hyd = s.search_vector_and_retrieve(conn, [1, 2, 3], 5, q)
        keep = ~np.isin(ids, victims)
        w_ids, w_scores, _ = orc.search(rows[keep], ids[keep], q, 5, mode=orc.MODE_F32_V1)
        assert [it.id for _, it in hyd] == w_ids.tolist()
        assert np.array_equal(np.array([it.score for _, it in hyd], dtype=np.float32), w_scores)
        s.hidden.discard(victims[0])  # un-hiding is picked up by the next search
        assert s.search_vector([1, 2, 3], 5, q)[0].id == victims[0]
        s.filter_hidden = False
        assert [t.id for t in s.search_vector([1, 2, 3], 5, q)] == [t.id for t in top]
        # --like: the stored embedding of an item is the query; it is its own best hit
        like = s.embedding_of(int(ids[40]))
        assert np.array_equal(like, rows[40]) and s.embedding_of(999_999) is None
        assert s.search_vector([1, 2, 3], 3, like)[0].id == int(ids[40])
    finally:
        s.close()


def test_bulk_decode_rejects_ragged_embeddings(pcv_lib, orc):
    """A row whose BLOB has another dimension is an error naming the row, not silent truncation."""
    from perceive_b200 import searcher
    conn, _ = make_db(orc, n=30)
    conn.execute("UPDATE item_embeddings SET embedding = ? WHERE item_id = 1004 AND model_id = 7", (b"\\x00" * 40,))
    with pytest.raises(ValueError, match="embedding 4|inconsistent"):
        searcher._load_rows(conn, 7, 0, [1, 2, 3])


def _file_db(orc, tmp_path, n=600):
    """The same fixture as make_db, written to a database FILE (the native reader opens a path)."""
    conn, live = make_db(orc, n=n)
    path = tmp_path / "perceive.db"
    disk = sqlite3.connect(path)
    conn.backup(disk)
    disk.close()
    return path, conn, live


def test_native_sqlite_loader_matches_the_python_loader(pcv_lib, orc, tmp_path):
    """pcv_rowset_from_sqlite (search.rs:87-113 in C++, libsqlite3 through dlopen) returns the same
    rows, ids and source ids as the reference SQL run through Python's sqlite3 module."""
    from perceive_b200 import searcher
    path, conn, live = _file_db(orc, tmp_path)
    for model, flt in ((7, [1, 2, 3]), (7, [1, 3]), (7, [2]), (7, []), (7, None), (3, [1, 2, 3]), (99, [1, 2, 3])):
        n_rows, n_ids, n_srcs, n_dim = searcher._load_rows_native(path, model, 0, flt)
        p_rows, p_ids, p_srcs, p_dim = searcher._load_rows(conn, model, 0, [1, 2, 3] if flt is None else flt)
        assert n_dim == p_dim
        on, op = np.argsort(n_ids), np.argsort(p_ids)  # neither side promises an order (no ORDER BY)
        assert np.array_equal(n_ids[on], p_ids[op]) and np.array_equal(n_srcs[on], p_srcs[op])
        assert np.array_equal(n_rows[on], p_rows[op])
    assert set(searcher._load_rows_native(path, 7, 0, [1, 2, 3])[1].tolist()) == set(live)


def test_native_sqlite_loader_errors(pcv_lib, orc, tmp_path):
    import perceive_b200 as pb
    from perceive_b200 import searcher
    with pytest.raises(pb.PcvError) as e:  # no such file: opened read-only, never created
        searcher._load_rows_native(tmp_path / "missing.db", 7, 0, [1])
    assert e.value.code == 1 and not (tmp_path / "missing.db").exists()
    other = tmp_path / "other.db"
    c = sqlite3.connect(other)
    c.execute("CREATE TABLE t (x)")
    c.commit()
    c.close()
    with pytest.raises(pb.PcvError) as e:  # a database without the reference schema
        searcher._load_rows_native(other, 7, 0, [1])
    assert e.value.code == 1 and "schema" in e.value.message
    path, conn, _ = _file_db(orc, tmp_path, n=30)
    disk = sqlite3.connect(path)
    disk.execute("UPDATE item_embeddings SET embedding = ? WHERE item_id = 1004 AND model_id = 7", (b"\x00" * 40,))
    disk.commit()
    with pytest.raises(pb.PcvError) as e:  # ragged rows name the item
        searcher._load_rows_native(path, 7, 0, [1, 2, 3])
    assert e.value.code == 1 and "1004" in e.value.message
    disk.execute("UPDATE item_embeddings SET embedding = ? WHERE item_id = 1004 AND model_id = 7", (b"\x00" * 41,))
    disk.commit()
    disk.close()
    with pytest.raises(pb.PcvError) as e:  # the reference's chunks_exact(4) would panic on chunk[3]
        searcher._load_rows_native(path, 7, 0, [1, 2, 3])
    assert "1004" in e.value.message


@pytest.mark.gpu
def test_searcher_build_from_a_database_file(pcv_lib, orc, tmp_path):
    """Searcher.build / rebuild_source given a path use the native reader end to end."""
    import perceive_b200 as pb
    path, conn, live = _file_db(orc, tmp_path)
    ids = np.array(sorted(live), dtype=np.int64)
    rows = np.stack([live[int(i)][1] for i in ids])
    srcs = np.array([live[int(i)][0] for i in ids], dtype=np.int64)
    q = orc.synth_rows(12, 0, 0, 1, DIM)[0]
    s = pb.Searcher.build(path, 7, 0)
    try:
        for flt in ([1, 2, 3], [2]):
            got = s.search_vector(flt, 10, q)
            w_ids, w_scores, _ = orc.search(rows, ids, q, 10, source_ids=srcs, sources=flt, mode=orc.MODE_F32_V1)
            assert [g.id for g in got] == w_ids.tolist()
            assert np.array_equal(np.array([g.score for g in got], dtype=np.float32), w_scores)
        disk = sqlite3.connect(path)
        top = s.search_vector([1, 2, 3], 3, q)
        disk.execute("UPDATE items SET hidden_at = 1 WHERE id = ?", (top[0].id,))
        disk.commit()
        disk.close()
        s.rebuild_source(path, live[top[0].id][0], 7, 0)
        assert s.search_vector([1, 2, 3], 3, q)[0].id == top[1].id
    finally:
        s.close()


def test_native_loader_reads_a_wal_database_next_to_a_live_writer(pcv_lib, orc, tmp_path):
    """The reference opens its database with journal=wal (crates/perceive-core/db.rs:94).  The native
    reader — a separate read-only connection — must see rows that are committed but still only in
    the write-ahead log, and must not see a writer's uncommitted rows."""
    import perceive_b200 as pb
    from perceive_b200 import searcher
    path = tmp_path / "wal.db"
    w = sqlite3.connect(path, isolation_level=None)
    assert w.execute("PRAGMA journal_mode=wal").fetchone()[0] == "wal"
    w.execute("PRAGMA wal_autocheckpoint=0")  # keep everything in the -wal file
    apply_reference_schema(w)
    w.execute("INSERT INTO sources (id, name, location, compare_strategy, status) VALUES (1,'s','/','mtime','ready')")
    vecs = orc.synth_rows(5, 0, 0, 6, DIM)

    def add(i):
        w.execute("INSERT INTO items (id, source_id, external_id, hash, content) VALUES (?,1,?,?,?)", (100 + i, str(i), "h", "x"))
        w.execute("INSERT INTO item_embeddings VALUES (7, 0, ?, 0, ?)", (100 + i, pb.serialize_embedding(vecs[i])))

    for i in range(4):
        add(i)  # autocommit: committed, not checkpointed
    assert (tmp_path / "wal.db-wal").stat().st_size > 0
    w.execute("BEGIN")
    add(4)  # uncommitted
    rows, ids, srcs, dim = searcher._load_rows_native(path, 7, 0, None)
    assert dim == DIM and sorted(ids.tolist()) == [100, 101, 102, 103]
    assert np.array_equal(rows[np.argsort(ids)], vecs[:4])
    w.execute("COMMIT")
    assert sorted(searcher._load_rows_native(path, 7, 0, [1])[1].tolist()) == [100, 101, 102, 103, 104]
    w.close()
