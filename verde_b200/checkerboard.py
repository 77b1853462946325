# Host glue derived from fatiando/verde (BSD 3-Clause): it restates the reference's estimator contract, error and
# warning texts so that this path is a drop-in.
#   Copyright (c) 2017 The Verde Developers.  Distributed under the terms of the BSD 3-Clause License.
#   SPDX-License-Identifier: BSD-3-Clause
#   This code is part of the Fatiando a Terra project (https://www.fatiando.org); see LICENSE-verde.txt.
"""
``CheckerBoard``: the synthetic field every BASELINE config samples
(verde/synthetic.py:16-197).  Host numpy; it exists here only as the input
generator for fixtures and benchmarks, not as a kernel target.
"""
import numpy as np

from .base import BaseGridder, check_data, get_instance_region, project_coordinates
from .coords import check_region, scatter_points


class CheckerBoard(BaseGridder):
    """
    ``amplitude * sin(2 pi e / w_east) * cos(2 pi n / w_north)`` on *region*;
    the wavelengths default to half the region's extent.
    """

    def __init__(self, amplitude=1000, region=(0, 5000, -5000, 0), w_east=None, w_north=None):
        super().__init__()
        self.amplitude = amplitude
        self.w_east = w_east
        self.w_north = w_north
        self.region = region

    @property
    def w_east_(self):
        "East wavelength actually used"
        return (self.region[1] - self.region[0]) / 2 if self.w_east is None else self.w_east

    @property
    def w_north_(self):
        "North wavelength actually used"
        return (self.region[3] - self.region[2]) / 2 if self.w_north is None else self.w_north

    @property
    def region_(self):
        "Lets the inherited grid/scatter/profile find a default region"
        check_region(self.region)
        return self.region

    def predict(self, coordinates):
        easting, northing = coordinates[:2]
        return (
            self.amplitude
            * np.sin((2 * np.pi / self.w_east_) * easting)
            * np.cos((2 * np.pi / self.w_north_) * northing)
        )

    def scatter(self, region=None, size=300, random_state=0, dims=None, data_names=None,
                projection=None, **kwargs):
        "Random sample of the field as a DataFrame (not deprecated for the synthetic)."
        dims = self._get_dims(dims)
        region = get_instance_region(self, region)
        coordinates = scatter_points(region, size, random_state=random_state, **kwargs)
        if projection is None:
            data = check_data(self.predict(coordinates))
        else:
            data = check_data(self.predict(project_coordinates(coordinates, projection)))
        return self._table(coordinates, data, dims, data_names)
