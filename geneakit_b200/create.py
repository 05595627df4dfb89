"""gen.genealogy (reference geneakit/create.py:6-80): DataFrame or file -> Pedigree."""
import os

import pandas as pd

from . import cgeneakit


def genealogy(input, **kwargs):
    sorted = kwargs.get('sorted', False)
    if isinstance(input, pd.DataFrame):
        dataframe = input
        ids = dataframe.iloc[:, 0].values
        father_ids = dataframe.iloc[:, 1].values
        mother_ids = dataframe.iloc[:, 2].values
        sexes = dataframe.iloc[:, 3].values
        sexes = [1 if sex == 'M' else 2 if sex == 'F' else 0 if sex == 'U' else sex for sex in sexes]
        return cgeneakit.load_pedigree_from_vectors(ids, father_ids, mother_ids, sexes, sorted)
    elif isinstance(input, str):
        if not os.path.exists(input):
            raise FileNotFoundError('File not found: ' + input)
        return cgeneakit.load_pedigree_from_file(input, sorted)
    raise ValueError('genealogy() takes a DataFrame or a file path')
