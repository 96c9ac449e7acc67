"""The reference's two plain-text inputs: list_with_focals.txt (utils.h:120-182) and the similarity matrix
(imagesimilarity_graph.h:108-171)."""
import numpy as np
import pytest

from pose_graph_initialization_b200 import builder as B
from pose_graph_initialization_b200 import scene as S


def test_image_list_with_focals(tmp_path):
    p = tmp_path / "list_with_focals.txt"
    p.write_text("images/100_1234.jpg 0 1520.5\nimages/abc.jpg\nimages/x y.jpg 1 800\n\nimages/z.jpg 0 notanumber\n")
    names, focal = S.load_1dsfm_image_list(str(p))
    assert names == ["100_1234.jpg", "abc.jpg", "x", "", "z.jpg"]  # first 7 characters dropped; tokens split on blanks
    assert np.array_equal(focal, [1520.5, 0.0, 1.0, 0.0, 0.0])     # third token (std::atof); missing -> 0.0 here


def test_similarity_matrix_round_trip_feeds_the_queue(tmp_path):
    sc = S.make_scene(n_views=7, n_corr=60, seed=4, n_points=200)
    p = tmp_path / "sim.txt"
    S.save_similarity_matrix(str(p), sc["sim"])
    sim = S.load_similarity_matrix(str(p), 7)
    assert np.array_equal(sim, sc["sim"])
    sc2 = dict(sc, sim=sim)
    q1 = B.HostBuilder(sc, host_threads=1, similarity_threshold=0.0).queue_pairs()
    q2 = B.HostBuilder(sc2, host_threads=1, similarity_threshold=0.0).queue_pairs()
    assert np.array_equal(q1, q2) and len(q1) == 21
    with pytest.raises(ValueError):
        S.load_similarity_matrix(str(p), 8)
    (tmp_path / "bad.txt").write_text("1 0.5\n0.5\n")
    with pytest.raises(ValueError):
        S.load_similarity_matrix(str(tmp_path / "bad.txt"), 2)
