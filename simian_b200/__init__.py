"""simian-b200: a B200-native conservative PDES engine behind Simian's
event-oriented service API (Simian / addEntity / schedService / reqService /
attachService / run).  See DESIGN.md."""
__version__ = "0.1.0"
