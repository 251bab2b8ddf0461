"""``Generator`` wrapper with the reference's surface (``bem/datasets/Data.py:17-89``) for the two
distributions on the hot path: 'skewed_levy' and 'sas'.  The toy 2-D datasets of the reference
(gmm_grid, swiss_roll, ...) are data loading, not the sampling path, and are out of scope."""
from inspect import signature

from .Distributions import gen_sas, gen_skewed_levy


class Generator:
    available_distributions = ["skewed_levy", "sas"]

    def __init__(self, operation, transform=None, *args, **kwargs):
        self.transform = (lambda x: x) if transform is None else transform
        self.kwargs = kwargs
        self.args = args
        self.samples = None
        table = {"skewed_levy": gen_skewed_levy, "sas": gen_sas}
        try:
            self.generator = table[operation]
        except KeyError:
            raise Exception("Unknown distribution to sample from. "
                            "Available distributions: {}".format(list(table.keys())))

    def setTransform(self, transform):
        self.transform = transform

    def setParams(self, *args, **kwargs):
        """Data.py:60-65: None positional entries keep the stored value; kwargs are merged."""
        if args == () and kwargs == {}:
            raise Exception("Given void parameters")
        self.args = tuple(map(lambda x, y: y if y is not None else x, self.args, args))
        self.kwargs.update(kwargs)

    def getSignature(self):
        return signature(self.generator)

    def generate(self, *args, **kwargs):
        """Data.py:75-83."""
        tmp_kwargs = self.kwargs | kwargs
        if args == () and kwargs == {} and self.kwargs == {}:
            raise Exception("No parameters for data generation")
        use_args = self.args if args == () else args
        self.samples = self.transform(self.generator(*use_args, **tmp_kwargs))
        return self.samples

    def __len__(self):
        return self.samples.size()

    def __getitem__(self, idx):
        return self.samples[idx]
