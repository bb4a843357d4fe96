"""The reference's own CubeList tests (lib/mpdaf/obj/tests/test_cubelist.py) on mpdaf_b200.CubeList.

Same fixtures (three 5x4x3 cubes arange(60)*{0,1,5}, float32 on disk, var = 1), same
assertions: exact data, expmap == 3, header propagation, scales/offsets.
"""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from mpdaf_b200 import CubeList
from mpdaf_b200 import fits as mfits


class TestCubeList:
    shape = (5, 4, 3)
    ncubes = 3
    cubevals = np.array([0, 1, 5])
    scalelist = [0.5, 1, 1.5]
    offsetlist = [1.5, 2, 2.5]

    @pytest.fixture(autouse=True)
    def _cubes(self, tmp_path):
        self.arr = np.arange(np.prod(self.shape)).reshape(self.shape)
        self.scaledcube = np.mean([(self.arr * val + self.offsetlist[k]) * self.scalelist[k]
                                   for k, val in enumerate(self.cubevals)], axis=0)
        self.combined_cube = np.mean([(self.arr * i) for i in self.cubevals], axis=0)
        self.cubenames = []
        for i in self.cubevals:
            filename = os.path.join(str(tmp_path), 'cube-%d.fits' % i)
            mfits.write_cube(filename, (self.arr * i).astype(np.float32), np.ones(self.shape, dtype=np.float32),
                             primary_header={'CUBEIDX': int(i), 'OBJECT': 'OBJECT %d' % i, 'EXPTIME': 100})
            self.cubenames.append(filename)
        self.expmap = np.full(self.shape, self.ncubes, dtype=int)

    def assert_header(self, cube):
        assert cube.primary_header['FOO'] == 'BAR'
        assert 'CUBEIDX' not in cube.primary_header
        assert cube.primary_header['OBJECT'] == 'OBJECT 0'
        assert cube.data_header['OBJECT'] == 'OBJECT 0'
        assert cube.primary_header['EXPTIME'] == 100 * self.ncubes

    def test_checks(self, tmp_path):
        other = os.path.join(str(tmp_path), 'cube-tmp.fits')
        mfits.write_cube(other, np.zeros((3, 2, 1), dtype=np.float32))
        clist = CubeList(self.cubenames[:1] + [other])
        assert clist.check_dim() is False
        assert CubeList(self.cubenames).check_dim() is True

    @pytest.mark.gpu
    def test_median(self):
        clist = CubeList(self.cubenames)
        cube, expmap, stat_pix = clist.median(header={'FOO': 'BAR'})
        self.assert_header(cube)
        assert_array_equal(cube.data, self.arr)
        assert_array_equal(expmap.data, self.expmap)
        assert stat_pix.colnames == ['FILENAME', 'NPIX_NAN']
        assert_array_equal(stat_pix['NPIX_NAN'], 0)

    @pytest.mark.gpu
    def test_combine(self):
        clist = CubeList(self.cubenames)
        cube, expmap, stat_pix = clist.combine(header={'FOO': 'BAR'})
        cube2, expmap2, stat_pix2 = clist.combine(header={'FOO': 'BAR'}, mad=True)
        assert_array_equal(cube.data, cube2.data)
        assert_array_equal(expmap.data, expmap2.data)

        self.assert_header(cube)
        assert_array_equal(cube.data, self.combined_cube)
        assert_array_equal(expmap.data, self.expmap)
        assert stat_pix.colnames == ['FILENAME', 'NPIX_NAN', 'NPIX_REJECTED']
        assert_array_equal(stat_pix['NPIX_REJECTED'], 0)
        assert cube.primary_header['NFILES'][0] == 3
        assert cube.primary_header['HIERARCH MPDAF METH1 ID'][0] == 'obj.cubelist.merging'

        cube = clist.combine(nclip=(5., 5.), var='stat_mean')[0]
        assert_array_equal(cube.data, self.combined_cube)

    @pytest.mark.gpu
    def test_combine_scale(self):
        clist = CubeList(self.cubenames, scalelist=self.scalelist, offsetlist=self.offsetlist)
        cube, expmap, stat_pix = clist.combine(header={'FOO': 'BAR'})
        # the reference's test asserts equality; the pair kernel's mean is pivot + sum(w - pivot) / n, a few ulp from
        # the exposure-order sum of w (contract: 1e-6 relative)
        assert_allclose(cube.data, self.scaledcube, rtol=1e-14, atol=0)

    @pytest.mark.gpu
    def test_pymedian(self):   # test_cubelist.py:132-138
        clist = CubeList(self.cubenames)
        cube, expmap, stat_pix = clist.pymedian(header={'FOO': 'BAR'})
        self.assert_header(cube)
        assert_array_equal(cube.data, self.arr)
        assert_array_equal(expmap.data, self.expmap)

    @pytest.mark.gpu
    def test_pycombine(self):   # test_cubelist.py:156-170
        clist = CubeList(self.cubenames)
        cube, expmap, stat_pix, rejmap = clist.pycombine(header={'FOO': 'BAR'})
        cube2, expmap2, stat_pix2, rejmap2 = clist.pycombine(header={'FOO': 'BAR'}, mad=True)
        assert_array_equal(cube.data, cube2.data)
        assert_array_equal(expmap.data, expmap2.data)
        self.assert_header(cube)
        assert_array_equal(cube.data, self.combined_cube)
        assert_array_equal(expmap.data, self.expmap)
        assert_array_equal(rejmap.data, 0)
        cube = clist.pycombine(nclip=(5., 5.), var='stat_mean')[0]
        assert_array_equal(cube.data, self.combined_cube)

    @pytest.mark.gpu
    def test_pycombine_scale(self):   # test_cubelist.py:179-184
        clist = CubeList(self.cubenames, scalelist=self.scalelist, offsetlist=self.offsetlist)
        cube2, expmap2, _, _ = clist.pycombine(header={'FOO': 'BAR'})
        assert_allclose(cube2.data, self.scaledcube, rtol=1e-14, atol=0)

    @pytest.mark.gpu
    def test_mosaic_combine(self):   # test_cubelist.py:186-197
        from mpdaf_b200 import CubeMosaic
        clist = CubeMosaic(self.cubenames, self.cubenames[0])
        cube, expmap, stat_pix, rejmap = clist.pycombine(header={'FOO': 'BAR'})
        self.assert_header(cube)
        assert_array_equal(cube.data, self.combined_cube)
        assert_array_equal(expmap.data, self.expmap)
        cube2, expmap2, _, _ = clist.pycombine(header={'FOO': 'BAR'}, mad=True)
        assert_array_equal(cube.data, cube2.data)
        assert_array_equal(expmap.data, expmap2.data)
        with pytest.raises(NotImplementedError):
            clist.combine()

    @pytest.mark.gpu
    def test_mosaic_offsets_and_rejmap(self, tmp_path):
        """Cubes placed at different CRPIX offsets inside a larger output grid; rejmap counts."""
        from mpdaf_b200 import CubeMosaic
        import oracle
        rng = np.random.default_rng(5)
        out_shape, sub = (6, 9, 11), (6, 5, 6)
        outfile = os.path.join(str(tmp_path), 'out.fits')
        mfits.write_cube(outfile, np.zeros(out_shape, dtype=np.float32), data_header={'CRPIX1': 6.0, 'CRPIX2': 5.0})
        offsets = [(0, 0), (4, 5), (2, 3), (1, 4)]
        names, padded, pvar = [], [], []
        for k, (oy, ox) in enumerate(offsets):
            d = (10 + rng.standard_normal(sub)).astype(np.float32)
            d[rng.random(sub) < 0.1] += 30
            v = rng.uniform(0.5, 2, sub).astype(np.float32)
            name = os.path.join(str(tmp_path), 'part-%d.fits' % k)
            mfits.write_cube(name, d, v, data_header={'CRPIX1': 6.0 - ox, 'CRPIX2': 5.0 - oy})
            names.append(name)
            full = np.full(out_shape, np.nan, dtype=np.float32)
            fv = np.full(out_shape, np.nan, dtype=np.float32)
            full[:, oy:oy + sub[1], ox:ox + sub[2]] = d
            fv[:, oy:oy + sub[1], ox:ox + sub[2]] = v
            padded.append(full)
            pvar.append(fv)
        cube, expmap, statpix, rejmap = CubeMosaic(names, outfile).pycombine(nclip=2.0)
        want = oracle.combine_port(padded, pvar, nclip=2.0)
        assert cube.data.shape == out_shape
        assert_array_equal(expmap.data.ravel(), want["expmap"])
        np.testing.assert_allclose(cube.data.ravel(), want["data"], rtol=1e-6, equal_nan=True)
        nvalid = sum((~np.isnan(p)).astype(int) for p in padded)
        assert_array_equal(rejmap.data, nvalid - expmap.data)
        assert (rejmap.data > 0).any() and (expmap.data == 0).any()

    @pytest.mark.gpu
    def test_from_arrays_and_write(self, tmp_path):
        arrays = [np.asarray(mfits.read_image(f, 'DATA'), dtype=np.float32) for f in self.cubenames]
        var = [np.ones(self.shape, dtype=np.float32)] * 3
        clist = CubeList.from_arrays(arrays, var)
        cube, expmap, stat_pix = clist.combine()
        assert_array_equal(cube.data, self.combined_cube)
        assert_array_equal(cube.var, np.full(self.shape, 3.0 / 9.0))
        out = os.path.join(str(tmp_path), 'combined.fits')
        cube.write(out)
        assert_array_equal(np.asarray(mfits.read_image(out, 'DATA')), self.combined_cube.astype(np.float32))
        assert mfits.read_hdus(out)[0][0]['NFILES'] == 3

    @pytest.mark.gpu
    def test_device_stack(self):
        from mpdaf_b200 import synth
        shape = (7, 6, 9)
        stack, _, _ = synth.device_stack(5, shape, seed=3)
        data, var = synth.host_stack(5, shape, seed=3)
        import oracle
        want = oracle.combine_port(data, var, nclip=3.0)
        cube, expmap, statpix = CubeList.from_arrays(stack).combine(nclip=3.0)
        assert_array_equal(expmap.data.ravel(), want["expmap"])
        np.testing.assert_allclose(cube.data.ravel(), want["data"], rtol=1e-6, equal_nan=True)
        assert_array_equal(statpix['NPIX_NAN'], int(np.prod(shape)) - want["valid_pix"])


@pytest.mark.gpu
def test_expnb_histogram_on_device():
    """_compute_expnb (cubelist.py:69-74) from the device-side exposure-map histogram."""
    from mpdaf_b200 import synth
    from mpdaf_b200.cubelist import _compute_expnb, _expnb_from_device
    data, var = synth.host_stack(7, (5, 12, 14), seed=9)
    clist = CubeList.from_arrays(data, var)
    cube, expmap, _ = clist.combine(nclip=2.0)
    assert _expnb_from_device(7) == _compute_expnb(expmap.data)
    from mpdaf_b200.ctools import ctools
    hist = np.zeros(8, dtype=np.int64)
    assert ctools.mpdaf_b200_last_expmap_histogram(hist, 8) == 0
    assert_array_equal(hist, np.bincount(expmap.data.ravel(), minlength=8))
    cube, expmap, _ = clist.median()
    assert _expnb_from_device(7) == _compute_expnb(expmap.data)
