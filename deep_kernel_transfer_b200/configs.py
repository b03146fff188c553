"""Module-level configuration with the reference's names (configs.py:1-7).  ``kernel_type`` is read at
import time by methods/DKT.py and methods/DKT_regression.py, exactly like the reference."""
save_dir = './save/'
data_dir = {}
data_dir['CUB'] = './filelists/CUB/'
data_dir['miniImagenet'] = './filelists/miniImagenet/'
data_dir['omniglot'] = './filelists/omniglot/'
data_dir['emnist'] = './filelists/emnist/'
kernel_type = 'bncossim'  # linear, rbf, spectral (regression only), matern, poli1, poli2, cossim, bncossim
