"""API-compat patches that let the reference's 2020-era host code (numpy 1.17 / scipy 1.3 /
sklearn 0.23) run unmodified on this image (numpy 2.3 / scipy 1.16+ / sklearn 1.7+).  None of them
changes numerics: `.A` is toarray(); `n_alphas=N` and `alphas=N` select the same geometric grid;
`normalize=False` was the no-op default.  Used only by baseline/run_pipeline.py."""


def apply():
    import scipy.sparse as sp
    for name in ("csr_matrix", "csc_matrix", "coo_matrix", "csr_array", "csc_array", "coo_array"):
        cls = getattr(sp, name, None)
        if cls is not None and not hasattr(cls, "A"):
            try:
                cls.A = property(lambda self: self.toarray())
            except Exception:
                pass
    import inspect
    import sklearn.linear_model as lm
    import sklearn.linear_model._coordinate_descent as cd

    en_params = inspect.signature(cd.ElasticNet.__init__).parameters
    cv_params = inspect.signature(cd.ElasticNetCV.__init__).parameters
    real_en, real_cv = cd.ElasticNet, cd.ElasticNetCV

    def ElasticNet(*a, normalize=False, **kw):          # keyword translation only; returns the stock estimator
        return real_en(*a, **kw)

    def ElasticNetCV(*a, normalize=False, n_alphas=None, **kw):
        if n_alphas is not None:
            if "n_alphas" in cv_params and cv_params["n_alphas"].default != "deprecated":
                kw["n_alphas"] = n_alphas
            else:
                kw["alphas"] = n_alphas                 # sklearn >= 1.7: an int `alphas` is the grid size
        return real_cv(*a, **kw)

    if "normalize" not in en_params:
        lm.ElasticNet = ElasticNet
    if "normalize" not in cv_params:
        lm.ElasticNetCV = ElasticNetCV
