"""GDAL-free core of the tile driver (dsen2_b200/s2_tiles_supres.py) against values worked out by hand from the reference's
formulas (testing/s2_tiles_supres.py:127-140,150-158,174-193,221-303,311-329,385-416) on synthetic products."""
import numpy as np
import pytest

from dsen2_b200 import s2_tiles_supres as st

D10 = ["B4, central wavelength 665 nm", "B3, central wavelength 560 nm", "B2, central wavelength 490 nm",
       "B8, central wavelength 842 nm"]
D20 = ["B5, central wavelength 705 nm", "B6, central wavelength 740 nm", "B7, central wavelength 783 nm",
       "B8A, central wavelength 865 nm", "B11, central wavelength 1610 nm", "B12, central wavelength 2190 nm"]
D60 = ["B1, central wavelength 443 nm", "B9, central wavelength 945 nm", "B10, central wavelength 1375 nm"]


def test_roi_rounding_to_60m_pixel_boundaries():
    # :127-135 -- min rounds down to a multiple of 6, max+1 rounds down to a multiple of 6 (minus one)
    assert st.clamp_roi(100., 50., 17., 260., 10980, 10980) == (12, 48, 95, 257)
    assert st.clamp_roi(10970., -5., 20000., 30., 10980, 10980) == (10968, 0, 10979, 29)
    xmin, ymin, xmax, ymax = st.clamp_roi(3., 3., 4., 4., 600, 600)      # smaller than one 60 m pixel: empty region
    assert xmax < xmin and ymax < ymin
    assert st.full_roi(10980, 10980) == (0, 0, 10979, 10979)             # whole raster is NOT rounded (:136-140)
    for roi in (st.clamp_roi(100., 50., 17., 260., 10980, 10980), st.clamp_roi(0., 0., 599., 599., 600, 600)):
        assert (roi[2] - roi[0] + 1) % 6 == 0 and (roi[3] - roi[1] + 1) % 6 == 0 and roi[0] % 6 == 0 and roi[1] % 6 == 0


def test_projected_to_pixel_inverts_the_geotransform():
    geot = (300000., 10., 0., 5000040., 0., -10.)
    assert st.projected_to_pixel(geot, 300105., 5000005.) == (10, 3)
    rot = (1000., 8., 6., 2000., 6., -8.)                                  # rotated grid: x = 1000 + 8c + 6r, y = 2000 + 6c - 8r
    assert st.projected_to_pixel(rot, 1000. + 8 * 7 + 6 * 3 + 0.5, 2000. + 6 * 7 - 8 * 3 - 0.5) == (7, 3)


def test_utm_choice_prefers_largest_coverage_unless_named():
    cands = [("10m resolution, UTM 32N", (0, 0, 9, 9)), ("10m resolution, UTM 33N", (0, 0, 19, 19))]
    idx, utm, roi, areas = st.choose_utm(cands)
    assert (idx, utm, roi) == (1, "UTM 33N", (0, 0, 19, 19)) and areas == {"UTM 32N": 100, "UTM 33N": 400}
    idx, utm, roi, _ = st.choose_utm(cands, "UTM 32N")
    assert (idx, utm, roi) == (0, "UTM 32N", (0, 0, 9, 9))


def test_band_selection_and_descriptions():
    (n10, i10), (n20, i20), (n60, i60), desc = st.select_bands(D10, D20, D60, run_60=False)
    assert n10 == ['B4', 'B3', 'B2', 'B8'] and i10 == [0, 1, 2, 3]
    assert n20 == ['B5', 'B6', 'B7', 'B8A', 'B11', 'B12'] and n60 == [] and i60 == []
    assert desc['B4'] == 'B4 (665 nm)' and desc['B8A'] == 'B8A (865 nm)'
    (n10, _), (n20, _), (n60, i60), desc = st.select_bands(D10, D20, D60, run_60=True)
    assert n60 == ['B1', 'B9'] and i60 == [0, 1] and 'B10' not in desc           # B10 is never super-resolved
    assert st.get_band_short_name('B8A (865 nm)') == 'B8A' and st.get_band_short_name('B12') == 'B12'
    assert st.validate_description('foo, bar', 'ENVI') == 'foo bar' and st.validate_description('foo, bar') == 'foo, bar'


def test_read_windows_and_geotransform_shift():
    assert st.read_windows(12, 48, 95, 257) == ((12, 48, 84, 210), (6, 24, 42, 105), (2, 8, 14, 35))
    assert st.shift_geotransform((300000., 10., 0., 5000040., 0., -10.), 12, 48) == (300120., 10., 0., 4999560., 0., -10.)


def _product(tmp_path, H=120, W=180):
    rng = np.random.RandomState(0)
    path = str(tmp_path / 'product.npz')
    np.savez(path, data10=rng.randint(0, 9000, (H, W, 4)).astype(np.uint16),
             data20=rng.randint(0, 9000, (H // 2, W // 2, 6)).astype(np.uint16),
             data60=rng.randint(0, 9000, (H // 6, W // 6, 3)).astype(np.uint16),
             desc10=np.array(D10), desc20=np.array(D20), desc60=np.array(D60),
             geotransform=np.array([300000., 10., 0., 5000040., 0., -10.]), utm=np.array('UTM 32N'))
    return path


def test_cli_chains_60m_then_20m_and_writes_npz(tmp_path, monkeypatch, capsys):
    """The driver end to end on an .npz product with the network calls stubbed (host logic only): 60 m first, then 20 m
    (:332-342); output = [original 10 m bands,] SR 20 m bands, SR 60 m bands (:385-416) as {description: array}."""
    from dsen2_b200 import supres
    calls = []

    def fake20(d10, d20, deep=False, model=None):
        calls.append(('20', d10.shape, d20.shape))
        return np.repeat(np.repeat(d20.astype(np.float32), 2, 0), 2, 1) + 1

    def fake60(d10, d20, d60, deep=False, model=None):
        calls.append(('60', d10.shape, d20.shape, d60.shape))
        return np.repeat(np.repeat(d60.astype(np.float32), 6, 0), 6, 1) + 2
    monkeypatch.setattr(supres, 'DSen2_20', fake20)
    monkeypatch.setattr(supres, 'DSen2_60', fake60)
    path = _product(tmp_path)
    out = str(tmp_path / 'sr.npz')
    assert st.main([path, out, '--roi_x_y', '13,7,100,70', '--run_60', '--copy_original_bands', '--output_file_format', 'npz']) == 0
    # ROI (13,7)-(100,70) -> xmin 12, xmax int(101/6)*6-1 = 95, ymin 6, ymax int(71/6)*6-1 = 65: 84 x 60 pixels
    assert calls == [('60', (60, 84, 4), (30, 42, 6), (10, 14, 2)), ('20', (60, 84, 4), (30, 42, 6))]
    bands = np.load(out, allow_pickle=True)['bands'].item()
    names = list(bands)
    assert names == ['B4 (665 nm)', 'B3 (560 nm)', 'B2 (490 nm)', 'B8 (842 nm)'] + \
        ['SR' + d for d in ('B5 (705 nm)', 'B6 (740 nm)', 'B7 (783 nm)', 'B8A (865 nm)', 'B11 (1610 nm)', 'B12 (2190 nm)',
                            'B1 (443 nm)', 'B9 (945 nm)')]
    z = np.load(path)
    assert np.array_equal(bands['B4 (665 nm)'], z['data10'][6:66, 12:96, 0])
    assert np.array_equal(bands['SRB5 (705 nm)'], np.repeat(np.repeat(z['data20'][3:33, 6:48, 0].astype(np.float32), 2, 0), 2, 1) + 1)
    assert np.array_equal(bands['SRB9 (945 nm)'], np.repeat(np.repeat(z['data60'][1:11, 2:16, 1].astype(np.float32), 6, 0), 6, 1) + 2)
    text = capsys.readouterr().out
    assert 'Selected pixel region: xmin=12, ymin=6, xmax=95, ymax=65' in text and 'Image size: width=84 x height=60' in text
    # without --run_60 only the 20 m network runs and B1 / B9 are not in the output
    calls.clear()
    assert st.main([path, out, '--output_file_format', 'npz']) == 0
    assert [c[0] for c in calls] == ['20'] and calls[0][1] == (120, 180, 4)
    assert len(np.load(out, allow_pickle=True)['bands'].item()) == 6


def test_cli_listings_and_gdal_fallback(tmp_path, capsys, monkeypatch):
    from dsen2_b200 import supres
    monkeypatch.setattr(supres, 'DSen2_20', lambda d10, d20, deep=False, model=None: np.zeros(d10.shape[:2] + (6,), np.float32))
    path = _product(tmp_path)
    assert st.main([path, '--list_UTM']) == 0
    assert 'UTM 32N (21600)' in capsys.readouterr().out
    assert st.main([path, '--list_bands']) == 0
    assert 'Selected 20m bands: B5 B6 B7 B8A B11 B12' in capsys.readouterr().out
    assert st.main(['not-opened.zip', '--list_output_file_formats']) == 0     # exits before the input is touched (:64-79)
    assert capsys.readouterr().out.startswith('npz:')
    out = str(tmp_path / 'x.tif')
    assert st.main([path, out]) == 0                          # GTiff requested, no GDAL here: npz fallback like the reference
    assert "Writing to npz as a fallback" in capsys.readouterr().out
    assert len(np.load(out + '.npz', allow_pickle=True)['bands'].item()) == 6
    with pytest.raises(SystemExit):
        st.main([str(tmp_path / 'S2A.zip'), out])            # a real product needs GDAL: loud error, no silent path
