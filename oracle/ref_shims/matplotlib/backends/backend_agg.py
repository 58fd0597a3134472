class FigureCanvasAgg:
    pass
