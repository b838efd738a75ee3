from .plot_runtime_boxplot import make_boxplot

__all__ = ["make_boxplot"]
