from .results import (coordinate_transformations_dict, dumps_coordinate_transformations, export_results,  # noqa: F401
                      format_coordinate_transformations, frame_results_from_tensors)
