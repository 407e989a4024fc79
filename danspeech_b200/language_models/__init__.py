"""Language-model handles.

In the reference these factories download a KenLM binary into ``~/.danspeech/lms/`` and return its path
(danspeech/language_models/*.py, utils/data_utils.py:43-80).  There is no downloading here: a factory returns the
path of an already present file -- the ``.klm`` the reference would have cached (KenLM probing binaries are read
directly, csrc/lm_load.cu), else the ARPA text form ``<name>.arpa`` -- and otherwise says where it looked.
``CustomLanguageModel`` is the identity, as in the reference.
"""
import os


def CustomLanguageModel(path):
    """Identity, as danspeech/language_models/custom_lm.py:3-14."""
    return path


def _cached(stem, cache_dir):
    root = cache_dir if cache_dir is not None else os.path.join(os.path.expanduser("~"), ".danspeech", "lms")
    for ext in (".klm", ".arpa"):
        path = os.path.join(root, stem + ext)
        if os.path.isfile(path):
            return path
    raise FileNotFoundError("language model %s.klm (or .arpa) not found in %s; this package does not download models -- "
                            "place the file there or pass a path to Recognizer(lm=...)" % (stem, root))


def _factory(name, stem):
    def get(cache_dir=None):
        return _cached(stem, cache_dir)
    get.__name__ = get.__qualname__ = name
    get.__doc__ = "Path of the cached %s language model (``cache_dir`` defaults to ~/.danspeech/lms)." % stem
    return get


# factory name -> file stem, as published by the reference (language_models/<stem>.py)
_PUBLISHED = {"DSL3gram": "dsl_3gram", "DSL5gram": "dsl_5gram", "DSLWiki3gram": "dsl_wiki_3gram",
              "DSLWiki5gram": "dsl_wiki_5gram", "DSLWikiLeipzig3gram": "dsl_wiki_leipzig_3gram",
              "Wiki3gram": "wiki_3gram", "Wiki5gram": "wiki_5gram", "Folketinget3gram": "folketinget_3gram",
              "DSL3gramWithNames": "dsl_3gram_names"}
for _name, _stem in _PUBLISHED.items():
    globals()[_name] = _factory(_name, _stem)
__all__ = ["CustomLanguageModel"] + sorted(_PUBLISHED)
